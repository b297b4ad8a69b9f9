"""The header-only C++ shim (include/ractip_prob.hpp) that speaks RactIP's own VF/VI/VVF types."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import has_gpu

ROOT = Path(__file__).resolve().parent.parent


def _build(lib, tmp_path):
    exe = tmp_path / "shim_check"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "shim_check.cpp"),
           "-L", str(ROOT / "ractip_b200"), "-lractip_prob", f"-Wl,-rpath,{ROOT / 'ractip_b200'}", "-o", str(exe)]
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return exe


def test_shim_compiles_and_host_entry_points_work(lib, tmp_path):
    exe = _build(lib, tmp_path)
    r = subprocess.run([str(exe), "host"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    # the shim's shuffle equals the Python binding's
    from ractip_b200 import zscore_shuffles
    s1 = "GAAAGACGCGCAUUUGUUAUCAUCAUCCCUGAAUUCAGAGAUGAAAUUUUGGCCACUCACGAGUGGCCUUUU"
    s2 = "GCCAGGGGUGCUCGGCAUAAGCCGAAGAUAUCGG"
    assert ("shuffle0 " + zscore_shuffles(s1, s2, 3, 1)[0][0]) in r.stdout
    if not has_gpu():
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_shim_fills_ractip_members(lib, stage, bundled, tmp_path):
    exe = _build(lib, tmp_path)
    s = bundled["sequences"]["DIS"]
    r = subprocess.run([str(exe), "gpu", s, s], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    py = stage.solve_probabilities(s, s)
    line = [l for l in r.stdout.splitlines() if l.startswith("L1")][0].split()
    vals = dict(zip(line[0::2], line[1::2]))
    assert int(vals["L1"]) == len(s) and int(vals["offset1[1]"]) == py.offset1[1]
    assert abs(float(vals["sum_bp1"]) - float(py.bp1.astype(np.float64).sum())) < 1e-4
    assert abs(float(vals["sum_up1"]) - float(py.up1.astype(np.float64).sum())) < 1e-3
    assert abs(float(vals["sum_hp"]) - float(py.hp.astype(np.float64).sum())) < 1e-4


@pytest.mark.gpu
def test_shim_multi_gpu_batch_equals_single_gpu(lib, bundled, tmp_path):
    """rp_multi_*: the shuffle batch cut into one contiguous block per visible GPU (host threads), results written
    into the caller's one buffer -- bit-identical to the one-GPU run.  (On a one-GPU box: one block.)"""
    import torch
    exe = _build(lib, tmp_path)
    r = subprocess.run([str(exe), "multi", bundled["sequences"]["MicA"], bundled["sequences"]["ompA"], "37"],
                       stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert f"devices {torch.cuda.device_count()} pairs 37" in r.stdout and "multi == single" in r.stdout
