"""bench.py's output contract, checked on the CPU through the reference arm (the oracle port):
stdout carries exactly ONE JSON line with the keys the driver reads; everything else goes to
stderr.  The CUDA arm needs a GPU and is exercised by the driver itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference_arm(*extra):
    command = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "small",
               "--steps", "2", "--warmup", "1", "--reference-walks", "128", *extra]
    return subprocess.run(command, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    done = run_reference_arm()
    assert done.returncode == 0, done.stderr[-2000:]
    lines = done.stdout.splitlines()
    assert len(lines) == 1, lines
    record = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in record, key
    assert record["impl"] == "reference" and record["metric"] == "skipgram_context_pairs_per_s"
    assert record["steps"] == 2 and record["warmup"] == 1 and record["value"] > 0
    assert record["vs_baseline"] is None and record["higher_is_better"] is True
    assert record["cpu_baseline"]["kind"] == "port" and record["cpu_baseline"]["cores"] >= 1
    assert record["e2e"] == {"value": record["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}
    assert "workload" in record["config"]


def test_library_chatter_on_stdout_is_sent_to_stderr():
    """claim_stdout() points file descriptor 1 at stderr: a child that writes to fd 1 directly
    (what NCCL's version banner does) cannot add a line to stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version x.y\\n'); print('python chatter'); bench.emit_result({'ok': 1})" % ROOT)
    done = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert done.returncode == 0, done.stderr
    assert done.stdout == '{"ok": 1}\n'
    assert "NCCL version x.y" in done.stderr and "python chatter" in done.stderr


def test_reference_arm_never_loads_the_product_library():
    """The CPU arm is the baseline the product is compared with: it may execute oracle/ (the port)
    but must not load libb2e.so, not even to generate the graph (oracle/graphgen.c does that)."""
    code = ("import sys, os; sys.path.insert(0, %r); os.chdir(%r); import bench; "
            "sys.argv = ['bench.py', '--impl', 'reference', '--config', 'small', '--steps', '1', '--warmup', '0', "
            "'--reference-walks', '64']; bench.main(); maps = open('/proc/self/maps').read(); "
            "sys.stderr.write('B2E=%%d ORACLE=%%d\\n' %% ('libb2e.so' in maps, 'liboracle.so' in maps))" % (ROOT, ROOT))
    done = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, B2E_CACHE="/tmp"))
    assert done.returncode == 0, done.stderr[-2000:]
    assert "B2E=0 ORACLE=1" in done.stderr
