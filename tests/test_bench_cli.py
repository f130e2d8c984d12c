"""bench.py's reference arm on a GPU-less box: `--impl reference` must run the unmodified reference from oracle/_ref on inputs made by the
host generator, load neither the product library nor CUDA, and print ONE JSON line with the contract's keys."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

import orclib

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not orclib.have_ref(), reason="prebuilt reference library not present")
def test_reference_arm_runs_without_gpu_or_product_library():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--cols", "64", "--steps", "2", "--warmup", "3",
                        "--ref-threads", "2"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "aggregator input 64Kbit-blocks/s" and d["unit"] == "blocks/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["config"]["reduced"] is True and d["config"]["workload"].startswith("c3")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb and cb["rows"]
    assert d["e2e"] == {"value": d["value"], "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm must not have mapped the product library or the CUDA runtime
    probe = subprocess.run([sys.executable, "-c",
                            "import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--cols', '16', '--steps', '1', '--warmup', '3', '--ref-threads', '1'];\n"
                            "import contextlib, io\n"
                            "buf = io.StringIO()\n"
                            "try:\n"
                            "    runpy.run_path(%r, run_name='__main__')\n"
                            "except SystemExit:\n"
                            "    pass\n"
                            "maps = open('/proc/self/maps').read()\n"
                            "sys.stderr.write('LOADED_PRODUCT=%%d LOADED_CUDART=%%d\\n' %% ('libbmb200.so' in maps, 'libcudart' in maps))\n" % str(ROOT / "bench.py")],
                           capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert "LOADED_PRODUCT=0" in probe.stderr, probe.stderr[-1500:]
