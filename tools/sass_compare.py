"""Development aid (no GPU needed): do two builds of libpgk carry the same machine code for the same kernels?

    python tools/sass_compare.py <csrc dir of build A> <csrc dir of build B>

Reads `cuobjdump -sass` of every object file, pairs kernels by name (the anonymous-namespace hash is dropped, and a
kernel of A that gained trailing template parameters in B is paired with B's instance whose extra parameters are all
0 -- how the opt-in flavours are added), and reports per object: byte-identical kernels, kernels that differ only in
register numbers / immediates (same opcode sequence), and kernels that really differ.  Used at the end of round 1 to
check that the opt-in paths written without a GPU left the default path's kernels exactly as the GPU had verified
them (build A = a worktree of the last commit that ran on the GPU)."""
import difflib
import glob
import os
import re
import subprocess
import sys


def kernels(obj):
    out, cur = {}, None
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    for line in txt.split('\n'):
        if 'Function :' in line:
            cur = re.sub(r'_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]{8}', 'ANON', line.split('Function :')[1].strip())
            out[cur] = []
        elif cur:
            m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);', line)
            if m:
                out[cur].append(m.group(1).strip())
    return out


def shape(ins):
    ins = re.sub(r'\bU?R\d+\b', 'R', ins)
    ins = re.sub(r'\bU?P\d\b', 'P', ins)
    return re.sub(r'0x[0-9a-f]+', 'H', ins)


def partner(name, b):
    if name in b:
        return name
    m = re.match(r'(.*?I(?:Li\d+E)+)(E.*)$', name)       # ...kernelILi8ELi1EE... -> try appending Li0E once, twice
    if m:
        for extra in ('Li0E', 'Li0ELi0E', 'Li0ELi0ELi0E'):
            cand = m.group(1) + extra + m.group(2)
            if cand in b:
                return cand
    return None


def main():
    da, db = sys.argv[1], sys.argv[2]
    worst = 0
    for oa in sorted(glob.glob(os.path.join(da, '*.o'))):
        ob = os.path.join(db, os.path.basename(oa))
        if not os.path.exists(ob):
            continue
        a, b = kernels(oa), kernels(ob)
        same = renamed = 0
        differ, missing = [], []
        for k, v in a.items():
            k2 = partner(k, b)
            if k2 is None:
                missing.append(k)
            elif b[k2] == v:
                same += 1
            else:
                sa, sb = [shape(i) for i in v], [shape(i) for i in b[k2]]
                if sa == sb:
                    renamed += 1
                else:
                    sm = difflib.SequenceMatcher(None, sa, sb, autojunk=False)
                    n = sum(max(o[2] - o[1], o[4] - o[3]) for o in sm.get_opcodes() if o[0] != 'equal')
                    differ.append((k, n))
        print('%-18s %3d kernels in A, %3d in B: %3d byte-identical, %3d same opcode sequence, %3d differ, %3d missing'
              % (os.path.basename(oa), len(a), len(b), same, renamed, len(differ), len(missing)))
        for k, n in differ:
            print('      differs in %d instructions: %s' % (n, k[-90:]))
        for k in missing:
            print('      missing in B: %s' % k[-90:])
        worst = max(worst, len(differ) + len(missing))
    sys.exit(1 if worst else 0)


if __name__ == '__main__':
    main()
