"""CPU checks of the measurement plumbing: bench.py's algorithmic FLOP count against SURVEY.md 8(d) / Appendix B, the
JSON contract of the reference arm, and that the committed roofline traffic figure is what the committed ncu capture
says."""
import csv
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

# GFLOP per image per iteration, fade-in variant, 3-channel 1024^2 model
SURVEY_GFLOP = {1: 12.39, 2: 55.89, 3: 121.15, 4: 186.43, 5: 251.77, 6: 317.21, 7: 382.85, 8: 448.91}


@pytest.mark.parametrize('depth', sorted(SURVEY_GFLOP))
def test_algorithmic_flops_match_the_survey(depth):
    got = bench.flops_per_image(depth, 3, True) / 1e9
    assert got == pytest.approx(SURVEY_GFLOP[depth], rel=2e-3)


def test_fade_in_and_channels_change_the_count_as_the_survey_says():
    # without the fade-in, or with one image channel, the count is lower by a fraction of a percent
    for d in (4, 6, 8):
        full = bench.flops_per_image(d, 3, True)
        assert 0.0 < 1.0 - bench.flops_per_image(d, 3, False) / full <= 2e-3
        assert 0.0 < 1.0 - bench.flops_per_image(d, 1, True) / full <= 2.5e-3


def test_configs_are_the_baseline_configs():
    with open(os.path.join(ROOT, 'BASELINE.json')) as f:
        base = json.load(f)
    assert 'images/sec' in base['metric']
    c = bench.CONFIGS
    assert (c['c2']['depth'], c['c2']['alpha'], c['c2']['n'], c['c2']['precision']) == (4, 0.5, 128, 'fp32')
    assert (c['c3']['depth'], c['c3']['alpha'], c['c3']['n'], c['c3']['precision']) == (6, 1.0, 32, 'bf16')
    assert (c['c4']['depth'], c['c4']['alpha'], c['c4']['n'], c['c4']['precision']) == (8, 0.3, 4, 'bf16')
    assert (c['c5']['depth'], c['c5']['n'], c['c5']['res'], c['c5']['ch']) == (5, 64, 128, 1)
    assert (c['c1']['depth'], c['c1']['alpha'], c['c1']['n']) == (0, 1.0, 16)


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the unmodified reference (baseline/_ref, when __graft_entry__.build() installed it;
    else the oracle port) on the host cores, one JSON line with impl / cpu_baseline / e2e."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'c1',
                          '--steps', '1', '--warmup', '1'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/sec' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith('c1:')


@pytest.mark.parametrize('cfg,family', [('c2', 'conv_tc_kernel'), ('c4', 'conv_thin_kernel')])
def test_committed_traffic_figure_is_what_the_committed_capture_says(cfg, family):
    with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
        entry = json.load(f)[cfg][family]
    per = {}
    with open(os.path.join(ROOT, 'profiles', 'r2b_traffic_%s.csv' % cfg)) as f:
        rows = csv.DictReader([l for l in f if not l.startswith('==')])
        for row in rows:
            if family not in row['Kernel Name'] or not row['Metric Name'].startswith('dram__bytes'):
                continue
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[row['Metric Unit']]
            per[int(row['ID'])] = per.get(int(row['ID']), 0.0) + float(row['Metric Value'].replace(',', '')) * scale
    ids = sorted(per)
    ids = ids[len(ids) // 2:]
    assert len(ids) == entry['launches']
    assert sum(per[i] for i in ids) / len(ids) == pytest.approx(entry['bytes_per_launch'], rel=1e-9)


def test_roofline_assembly_from_launch_records():
    """bench.assemble_roofline is a pure function of the per-family launch records: feed it the figures of the
    committed end-of-round c2 run (gpurun_out/call8) and check the JSON it builds, including the tensor-pipe view
    (issued bf16 products: 6 per FLOP in forward passes, 3 in the gradient chains of the fp32-faithful mode)."""
    cfg = bench.CONFIGS['c2']
    ksteps, step_s = 3, 88.6e-3
    conv_ms, wgrad_ms = 58.08 * ksteps, 16.34 * ksteps
    conv_fl, wgrad_fl = 296.37e12 * conv_ms * 1e-3, 404.64e12 * wgrad_ms * 1e-3
    fam = {0: (conv_fl, 694.8e9 * conv_ms * 1e-3, conv_ms, 261), 1: (wgrad_fl, 721.5e9 * wgrad_ms * 1e-3, wgrad_ms, 57),
           2: (0.0, 0.0, 0.0, 0), 3: (1.0e9, 1.0e6, 0.2, 3), 4: (0.0, 0.0, 0.0, 0), 5: (0.0, 0.0, 0.0, 0)}
    prod = {0: 4.7 * conv_fl, 1: 3.0 * wgrad_fl, 2: 0.0, 3: 0.0, 4: 0.0, 5: 0.0}
    roof = bench.assemble_roofline('c2', cfg, fam, prod, ksteps, step_s)
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
        peak = json.load(f)['bf16_tflops_sustained']
    assert roof['bound'] == 'tensor' and roof['kernel'].startswith('conv_tc_kernel') and roof['unit'] == 'TFLOP/s'
    assert roof['peak'] == peak and roof['achieved'] == pytest.approx(296.37, rel=1e-6)
    assert roof['frac'] == pytest.approx(296.37 / peak, rel=1e-6)
    assert roof['tensor_pipe']['products_per_flop'] == pytest.approx(4.7)
    assert roof['tensor_pipe']['issued_tflops'] == pytest.approx(4.7 * 296.37, rel=1e-6)
    assert roof['tensor_pipe']['frac_of_peak'] == pytest.approx(4.7 * 296.37 / peak, rel=1e-6)
    assert roof['share_of_step'] == pytest.approx(58.08 / 88.6, rel=1e-6) and roof['launches_timed'] == 261
    assert set(roof['families']) == {'conv_tc_kernel', 'wgrad_tc_kernel', 'wgrad_simt_kernel'}
    assert roof['families']['wgrad_tc_kernel']['issued_tflops'] == pytest.approx(3 * 404.64, rel=1e-6)
    assert roof['traffic'] is not None and roof['algorithmic_bytes_per_launch'] > 0     # profiles/traffic.json, c2
    assert roof['step_algorithmic']['gflop_per_image'] == pytest.approx(186.41, rel=1e-3)
    assert 'note' in roof
    # a thin-layer family on top: the HBM roofline is reported, and a batch override drops the committed traffic figure
    fam[4] = (1.0e12, 3.0e9 * 100.0, 100.0e3 * 1e-3 * 1000, 10)
    prod[4] = 1.0e12
    r2 = bench.assemble_roofline('c4', bench.CONFIGS['c4'], fam, prod, ksteps, step_s, batch_overridden=True)
    assert r2['bound'] == 'hbm' and r2['unit'] == 'GB/s' and r2['traffic'] is None and 'note' not in r2
    json.dumps(roof), json.dumps(r2)
