"""In-tree build of every native artefact (no JIT cache: the built files travel with the repo snapshot).

    python flashattention.c_b200/build.py [--force] [--no-torch-ext] [--no-ref]

Artefacts
  flashattention.c_b200/libfa_b200.so        C-ABI library (include/fa_b200.h), sm_100a, static cudart
  flashattention.c_b200/flash_b200.so        torch extension `forward(Q,K,V,causal)` (links libfa_b200.so)
  flashattention.c_b200/harness/{fa_check,umma_probe,test}   torch-free harness binaries
  tests/harness/llmc_main                    the reference's llm.c harness (kernel 6) over the C symbols + oracle CPU loop
  compat/build/flash/flash.so                compat/src pre-built the way bench_flashattention.py:10 builds it
  oracle/_build/libfa_oracle.so              CPU restatement (checker only)
  oracle/_ref/*                              the reference itself, compiled from /root/reference when present
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HARNESS = PKG / "harness"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _run(cmd, **kw):
    print("[build]", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_lib(force=False) -> Path:
    out = PKG / "libfa_b200.so"
    srcs = [CSRC / "fa_api.cu", CSRC / "fa_fwd_sm100.cuh", CSRC / "fa_bwd_sm100.cuh", CSRC / "fa_simt.cuh", CSRC / "ptx.cuh", ROOT / "include" / "fa_b200.h"]
    if force or _stale(out, srcs):
        _run([NVCC, *NVCC_FLAGS, "--shared", "-o", out, CSRC / "fa_api.cu"])
    return out


def build_harness(force=False):
    lib = build_lib(force)
    outs = []
    for name, needs_lib in (("fa_check", True), ("test", True), ("umma_probe", False), ("mma_rate_probe", False)):
        src = HARNESS / f"{name}.cu"
        out = HARNESS / name
        deps = [src, CSRC / "ptx.cuh"] + ([lib] if needs_lib else [])
        if force or _stale(out, deps):
            cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-o", out, src]
            if needs_lib:
                cmd += [f"-L{PKG}", "-lfa_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/.."]
            _run(cmd)
        outs.append(out)
    return outs


def build_test_harness(force=False) -> Path:
    """tests/harness/llmc_main: the reference's llm.c harness (kernel-6 branch) over the exported C symbols, with the oracle's
    CPU loop as its checker — test infrastructure, which is why it lives under tests/ and not in the package."""
    lib = build_lib(force)
    build_oracle(force, with_ref=False)
    src = ROOT / "tests" / "harness" / "llmc_main.cu"
    out = ROOT / "tests" / "harness" / "llmc_main"
    orc = ROOT / "oracle" / "_build"
    if force or _stale(out, [src, lib, orc / "libfa_oracle.so"]):
        _run([NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-o", out, src, f"-L{PKG}", "-lfa_b200", f"-L{orc}", "-lfa_oracle",
              "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../flashattention.c_b200", "-Xlinker", "-rpath", "-Xlinker",
              "$ORIGIN/../../oracle/_build"])
    return out


def build_compat(force=False):
    """Pre-builds compat/src/{main.cpp,flashattention.cu} exactly as the reference's bench_flashattention.py:10 does
    (torch.utils.cpp_extension.load, name 'flash', cwd = compat/) into compat/build, so that running the script on the GPU
    box with TORCH_EXTENSIONS_DIR=compat/build finds an up-to-date build instead of paying the JIT."""
    compat = ROOT / "compat"
    so = compat / "build" / "flash" / "flash.so"
    srcs = [compat / "src" / "main.cpp", compat / "src" / "flashattention.cu", ROOT / "include" / "fa_b200.h"]
    if not (force or _stale(so, srcs)):
        return so
    env = dict(os.environ, TORCH_EXTENSIONS_DIR=str(compat / "build"), TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS="4")
    code = ("from torch.utils.cpp_extension import load; "
            "load(name='flash', sources=['src/main.cpp', 'src/flashattention.cu'], extra_cuda_cflags=['-O3'], verbose=True)")
    _run([sys.executable, "-c", code], cwd=str(compat), env=env)
    return so


def build_torch_ext(force=False) -> Path:
    """g++-only build of the pybind module (no device code in it)."""
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths

    lib = build_lib(force)
    out = PKG / "flash_b200.so"
    src = CSRC / "torch_binding.cpp"
    if not (force or _stale(out, [src, lib])):
        return out
    inc = [f"-I{p}" for p in include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    libdirs = [f"-L{p}" for p in library_paths("cuda")]
    torch_lib = Path(torch.__file__).parent / "lib"
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=flash_b200", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc, src, "-o", out, f"-L{PKG}", "-lfa_b200",
           *libdirs, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
    _run(cmd)
    return out


def build_oracle(force=False, with_ref=True):
    mk = ROOT / "oracle" / "Makefile"
    if not mk.exists():
        return
    targets = ["oracle"]
    if with_ref and Path("/root/reference/src/flashattention.cu").exists():
        targets.append("ref")
    _run(["make", f"-j{min(4, os.cpu_count() or 1)}", "-C", ROOT / "oracle", *targets] + (["-B"] if force else []))


def build_all(force=False, torch_ext=True, with_ref=True):
    build_lib(force)
    build_harness(force)
    if torch_ext:
        build_torch_ext(force)
    build_oracle(force, with_ref)
    build_test_harness(force)
    if torch_ext:
        build_compat(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, torch_ext="--no-torch-ext" not in sys.argv, with_ref="--no-ref" not in sys.argv)
