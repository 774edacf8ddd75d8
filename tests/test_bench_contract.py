"""bench.py contract on CPU: the reference arm prints ONE JSON line with the agreed keys; our arm refuses to run
without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--splats", "20000", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-6
    assert d["config"]["workload"] == "train_1m_1080p" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)


def test_algorithmic_bytes_one_pass_and_slab():
    """SURVEY.md 8(d) byte model: the one-pass formula, and the slab path's own passes when the library reports slabs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    st = dict(n_splats=1000, n_sorted=800, n_visible=600, n_tile_instances=5000, n_tiles=100, width=160, height=160,
              n_slabs=0, n_slabs_run=0, n_slab_entries=0, n_instances_kept=900)
    ab = bench.algorithmic_bytes(st)
    assert ab["sort"] == 20 * 1000 + 8 * 800 and ab["raster"] == 36 * 5000 + 4 * 160 * 160
    assert ab["total"] == ab["sort"] + ab["project"] + ab["bin"] + ab["raster"]
    st.update(n_slabs=3, n_slabs_run=2, n_slab_entries=700)
    sb = bench.algorithmic_bytes(st)
    assert sb["sort"] == 36 * 1000 and sb["project"] == 0
    assert sb["bin"] == 2 * 4 * 1000 + 700 * 86 + 80 * 900 + 8 * 100 * 2
    assert sb["total"] == sb["sort"] + sb["bin"] + sb["raster"]
