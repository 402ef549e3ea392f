"""bench.py's reference arm runs on CPU: check the one-JSON-line contract (keys, types, nothing else on stdout)
and the accounting helpers that need no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--ref-reads', '4', '--config', 'short'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'] == 'resquiggled raw samples/sec'
    assert line['unit'] == 'samples/s' and line['higher_is_better'] is True and line['value'] > 0
    assert line['steps'] == 1 and line['warmup'] == 0 and line['n_gpus'] == 1
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['vs_baseline'] is None and line['gpu_launches'] == 0


def test_algorithmic_bytes_match_the_survey_figure():
    sys.path.insert(0, ROOT)
    import bench
    # SURVEY.md 8d standard read: S = 89 000, Br = Bg = 10 000, A = 10 900, I = 1 200
    t = {'S': 89_000, 'Br1': 10_001, 'A': 10_900, 'Bg': 10_000, 'I': 1_200, 'S_reg': 31_000, 'R': 1}
    b = bench.algorithmic_bytes(t)
    assert b['path'] == 2 * 89_000 + 4 * 10_001 + 2 * 10_900 + 10_000 + 25 * 10_000 + 32
    assert abs(b['path'] / 89_000 - 5.6) < 0.1          # "about 5.5 B/sample"
    assert b['k_norm_stats'] == 2 * 89_000 + 32 and b['k_event_table'] == 2 * 89_000 + 4 * 10_001 + 24 * 10_000 + 32
    assert set(bench.STAGE_KERNELS) == set(bench.STAGES)
    path = os.path.join(ROOT, 'profiles', bench.TRAFFIC_FILE)
    if os.path.exists(path):  # the ncu capture bench.py quotes must be of the kernels it times
        traffic = json.load(open(path))['kernels']
        for names in bench.STAGE_KERNELS.values():
            for n in names:
                assert n in traffic and traffic[n]['warp_instructions'] > 0
