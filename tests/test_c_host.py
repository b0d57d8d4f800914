"""The C ABI from a host written in plain C (examples/c_host.c): no Python, no torch on that side.

CPU: the public header is valid C99 and the example compiles and links against the in-tree library.
GPU: the C program's results on a synthetic CT stack equal the Python shim's on the same stack."""

from __future__ import annotations

import os
import re
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "medical-image-enhancer_b200"
CUDA = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))


def _build(tmp_path: Path) -> Path:
    gcc = shutil.which("gcc")
    if gcc is None or not (CUDA / "include" / "cuda_runtime_api.h").exists():
        pytest.skip("needs gcc and the CUDA headers")
    exe = tmp_path / "c_host"
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), "-I", str(CUDA / "include"),
           str(ROOT / "examples" / "c_host.c"), "-o", str(exe), "-L", str(PKG), "-lmdimg_b200",
           "-L", str(CUDA / "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{CUDA / 'lib64'}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_c99_and_the_example_links(tmp_path):
    exe = _build(tmp_path)
    assert exe.exists()
    # the header alone, pedantically (no CUDA headers involved)
    src = tmp_path / "hdr.c"
    src.write_text('#include "mdimg_b200.h"\nint main(void) { mdimg_enhance_plan p = {0}; return (int)sizeof(p) == 0; }\n')
    res = subprocess.run([shutil.which("gcc"), "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I",
                          str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


@pytest.mark.gpu
def test_c_host_matches_the_python_shim(tmp_path, ops, synth):
    import torch
    from mdimg_b200.batch import process_stack
    exe = _build(tmp_path)
    n, size = 3, 128
    raw = np.stack([synth.ct_slice(7000 + z, z / 3, size=size) for z in range(n)])
    (tmp_path / "stack.u16").write_bytes(raw.tobytes())
    res = subprocess.run([str(exe), str(tmp_path / "stack.u16"), str(n), str(size), str(size), str(tmp_path / "enh.f32")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    enh_c = np.frombuffer((tmp_path / "enh.f32").read_bytes(), np.float32).reshape(n, size, size)
    py = process_stack(torch.from_numpy(raw.view(np.int16)).to(ops.device), synth.plan_full(), chunk=n, ops=ops)
    enh_py = py.enhanced.cpu().numpy()
    # the C host builds its Gaussian / bilateral weights with the C library's exp: a weight may differ from
    # numpy's in its last bit, which can move a pixel by an ulp or two
    assert float(np.abs(enh_c - enh_py).max()) <= 5e-7
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("slice ")]
    assert len(lines) == n
    for i, ln in enumerate(lines):
        kv = dict(re.findall(r"(\w+) (-?[\w.+-]+)", ln))
        mb, ma, val = py.metrics_before(i), py.metrics_after(i), py.validation(i)
        assert int(kv["flags"]) == (1 * bool(py.packed[i, 50]) | 2 * bool(py.packed[i, 51]) | 4 * bool(py.packed[i, 52]))
        assert int(kv["tv_iters"]) == int(py.tv_iterations[i])
        assert float(kv["sigma_before"]) == pytest.approx(mb["sigma"], rel=1e-9)
        assert float(kv["sigma_after"]) == pytest.approx(ma["sigma"], rel=1e-5)
        assert float(kv["entropy_after"]) == pytest.approx(ma["entropy"], rel=1e-5)
        assert float(kv["ssim"]) == pytest.approx(val["ssim"], rel=1e-6)
        assert float(kv["psnr"]) == pytest.approx(val["psnr"], rel=1e-6)
        assert float(kv["quality_improvement"]) == pytest.approx(val["quality_improvement"], rel=1e-4, abs=1e-6)
        assert int(kv["passes"]) == int(bool(val["passes"]))
        assert round(float(kv["score"]), 2) == pytest.approx(py.score(i)[0], abs=2e-2)
    assert int(re.search(r"launches (\d+)", res.stdout).group(1)) > 50
