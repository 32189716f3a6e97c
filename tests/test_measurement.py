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
    """bench.py --impl reference: the oracle port on the host cores, one JSON line with impl / cpu_baseline / e2e."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'c1',
                          '--steps', '1', '--warmup', '1'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/sec' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith('c1:')


def test_committed_traffic_figure_is_what_the_committed_capture_says():
    with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
        entry = json.load(f)['c2']['conv_tc_kernel']
    per = {}
    with open(os.path.join(ROOT, 'profiles', 'r1e_traffic_c2.csv')) as f:
        rows = csv.DictReader([l for l in f if not l.startswith('==')])
        for row in rows:
            if 'conv_tc_kernel' not in row['Kernel Name'] or not row['Metric Name'].startswith('dram__bytes'):
                continue
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[row['Metric Unit']]
            per[int(row['ID'])] = per.get(int(row['ID']), 0.0) + float(row['Metric Value'].replace(',', '')) * scale
    ids = sorted(per)
    ids = ids[len(ids) // 2:]
    assert len(ids) == entry['launches']
    assert sum(per[i] for i in ids) / len(ids) == pytest.approx(entry['bytes_per_launch'], rel=1e-9)
