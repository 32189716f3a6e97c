// Host check of the tiled weight re-layout (pggan-pytorch_b200/csrc/pgk_relayout.cuh): the two phases of each tiled
// kernel are run thread by thread (threads of a phase are independent; the __syncthreads() between the phases is the
// loop boundary) for every CTA of the grid the launcher would use, and every element of wf / wb / dw is compared with
// the one-thread-per-element mapping (weight_index) that the GPU-verified kernels use.  Built and run by
// tests/test_host.py with g++; no GPU involved.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../pggan-pytorch_b200/csrc/pgk_relayout.cuh"

static int check(int kind, int cin, int cin_stride, int cout, int ks, int accumulate) {
    const int taps = ks * ks;
    const long long nsrc = (long long)cout * cin_stride * taps, nop = (long long)cout * cin * taps;
    std::vector<float> w(nsrc), wf(nop, -1.f), wb(nop, -1.f), wf_ref(nop, -2.f), wb_ref(nop, -2.f);
    for (long long i = 0; i < nsrc; ++i) w[i] = (float)((i * 2654435761u) % 100003) / 100003.f - 0.5f;
    const float c = 0.37f;
    // reference: one element at a time
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int ky = 0; ky < ks; ++ky)
                for (int kx = 0; kx < ks; ++kx) {
                    long long fi, bi;
                    weight_index(kind, cin, cout, ks, co, ci, ky, kx, fi, bi);
                    const float v = c * w[(((long long)co * cin_stride + ci) * ks + ky) * ks + kx];
                    wf_ref[fi] = v, wb_ref[bi] = v;
                }
    // tiled: the launcher's grid, 256 threads per CTA, shared memory poisoned per CTA
    const int nthreads = 256, gx = (cin + relayout_tci(ks) - 1) / relayout_tci(ks), gy = (cout + kTileCo - 1) / kTileCo;
    std::vector<float> tile(relayout_tile_floats(ks));
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            const RelayoutTile t = relayout_tile(kind, cin, cin_stride, cout, ks, bx, by);
            if ((long long)kTileCo * t.row > (long long)tile.size()) return printf("tile overflow\n"), 1;
            for (auto& x : tile) x = NAN;
            for (int tid = 0; tid < nthreads; ++tid) relayout_load_w(tile.data(), w.data(), c, t, tid, nthreads);
            for (int tid = 0; tid < nthreads; ++tid) relayout_store_fb(tile.data(), wf.data(), wb.data(), t, tid, nthreads);
        }
    for (long long i = 0; i < nop; ++i)
        if (wf[i] != wf_ref[i] || wb[i] != wb_ref[i])
            return printf("prep mismatch kind %d cin %d/%d cout %d ks %d at %lld\n", kind, cin, cin_stride, cout, ks, i), 1;
    // gradient direction: dwp (wf layout) -> dw (PyTorch layout, stride cin_stride; the extra channels are not touched)
    std::vector<float> dwp(nop), dw(nsrc), dw_ref(nsrc);
    for (long long i = 0; i < nop; ++i) dwp[i] = (float)((i * 40503u) % 9973) / 9973.f;
    for (long long i = 0; i < nsrc; ++i) dw[i] = dw_ref[i] = 0.25f + (float)(i % 7);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int ky = 0; ky < ks; ++ky)
                for (int kx = 0; kx < ks; ++kx) {
                    long long fi, bi;
                    weight_index(kind, cin, cout, ks, co, ci, ky, kx, fi, bi);
                    const long long o = (((long long)co * cin_stride + ci) * ks + ky) * ks + kx;
                    const float v = c * dwp[fi];
                    dw_ref[o] = accumulate ? dw_ref[o] + v : v;
                }
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            const RelayoutTile t = relayout_tile(kind, cin, cin_stride, cout, ks, bx, by);
            for (auto& x : tile) x = NAN;
            for (int tid = 0; tid < nthreads; ++tid) relayout_load_dwp(tile.data(), dwp.data(), t, tid, nthreads);
            for (int tid = 0; tid < nthreads; ++tid) relayout_store_dw(tile.data(), dw.data(), c, accumulate, t, tid, nthreads);
        }
    for (long long i = 0; i < nsrc; ++i)
        if (dw[i] != dw_ref[i])
            return printf("unprep mismatch kind %d cin %d/%d cout %d ks %d acc %d at %lld\n", kind, cin, cin_stride, cout, ks, accumulate, i), 1;
    return 0;
}

int main() {
    int bad = 0, n = 0;
    const int shapes[][5] = {
        // kind, cin, cin_stride, cout, ks
        {PGK_W_CONV, 64, 64, 64, 3},   {PGK_W_CONV, 128, 128, 256, 3}, {PGK_W_CONV, 8, 8, 16, 3},
        {PGK_W_CONV, 16, 16, 8, 3},    {PGK_W_CONV, 32, 32, 64, 3},    {PGK_W_CONV, 48, 48, 40, 3},
        {PGK_W_CONV, 64, 65, 64, 3},   /* the stddev layer: one more stored input channel than is re-laid */
        {PGK_W_CONV, 3, 3, 16, 1},     {PGK_W_CONV, 64, 64, 3, 1},     {PGK_W_CONV, 100, 100, 33, 1},
        {PGK_W_GFIRST, 64, 64, 64, 4}, {PGK_W_GFIRST, 40, 40, 24, 4},  {PGK_W_DLAST, 64, 64, 64, 4},
        {PGK_W_DLAST, 72, 72, 48, 4},
    };
    for (const auto& s : shapes)
        for (int acc = 0; acc < 2; ++acc, ++n) bad += check(s[0], s[1], s[2], s[3], s[4], acc);
    printf("%d cases, %d bad\n", n, bad);
    return bad ? 1 : 0;
}
