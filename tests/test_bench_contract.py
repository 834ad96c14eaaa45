"""bench.py pieces that can be checked without a GPU: the algorithmic byte counts behind `roofline.achieved`
and the command-line contract."""
import importlib.util
import os
import subprocess
import sys

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_jh_algorithmic_bytes_match_the_survey():
    """B_JH = E2(3s+8) + E3(4s+8) + nHpl 18s + Pall 12s + Lall 3s + P 42s + L 12s -> fp64: ba_kitti_07
    20 504 008 B, ba_kitti_00 118 706 256 B; fp32 ba_kitti_00 61 597 592 B.  The landmark-pass kernel's share leaves out the
    Hpp/bp output (P 42s), which the pose pass writes."""
    b = _bench()
    k07 = dict(E2=20329, E3=74708, nhpl=94605, Pall=248, Lall=26127, numP=247, numL=26127)
    k00 = dict(E2=131233, E3=429883, nhpl=560658, Pall=1322, Lall=133383, numP=1321, numL=133383)
    assert b.jh_bytes(k07, 8)[1] == 20504008
    assert b.jh_bytes(k00, 8)[1] == 118706256
    assert b.jh_bytes(k00, 4)[1] == 61597592
    kernel, stage = b.jh_bytes(k00, 8)
    assert stage - kernel == 1321 * 42 * 8


def test_schur_algorithmic_bytes_match_the_survey():
    """B_S = nHpl 18s + L 12s + L 9s + nblk 36s + P 48s, ba_kitti_00 fp64 ~ 115.5 MB"""
    b = _bench()
    k00 = dict(nhpl=560658, numL=133383, nblk=41307, numP=1321)
    assert b.schur_bytes(k00, 8) == 560658 * 144 + 133383 * 168 + 41307 * 288 + 1321 * 384 == 115546776


def test_bench_refuses_to_run_without_a_gpu():
    """no CPU fallback: on a box without a CUDA device the engine arm exits with an error instead of measuring something else"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
