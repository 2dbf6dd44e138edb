"""In-tree build of the sm_100a extension ``tiny_deepspeed_b200/_C.so``.

* ``*.cu`` kernels are plain CUDA (no torch headers → seconds per file), compiled in parallel with
  ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` — sm_100a ONLY, no other arch, no
  fallback backend;
* ``bindings.cpp`` is the single torch/pybind translation unit, compiled with g++;
* the result is linked into ``tiny_deepspeed_b200/_C.so`` next to the sources so it travels with the
  tree to the GPU box (a JIT cache under ~/.cache would not).

``load()`` imports the module, rebuilding first when sources changed and a compiler is available;
on a machine with a GPU but no built extension it raises (no silent PyTorch fallback).
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
SO_PATH = PKG / "_C.so"
HASH_PATH = PKG / "_C.hash"
OBJ_DIR = HERE / "_build"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v", "-DNDEBUG"]


def _sources():
    cu = sorted(HERE.glob("*.cu"))
    cpp = sorted(HERE.glob("*.cpp"))
    hdr = sorted(list(HERE.glob("*.cuh")) + list(HERE.glob("*.h")))
    return cu, cpp, hdr


def source_hash() -> str:
    cu, cpp, hdr = _sources()
    h = hashlib.sha256()
    for f in cu + cpp + hdr + [Path(__file__)]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("CUDA_HOME", "") + "/bin/nvcc", "/usr/local/cuda/bin/nvcc", shutil.which("nvcc") or ""):
        if cand and os.path.isfile(cand):
            return cand
    return None


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.write(" ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr + "\n")
    if r.returncode != 0:
        raise RuntimeError(f"build failed:\n{' '.join(map(str, cmd))}\n{r.stdout}\n{r.stderr}")
    return r.stdout + r.stderr


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile everything for sm_100a and link ``_C.so``.  Returns the path of the library."""
    import torch
    from torch.utils import cpp_extension as ce

    digest = source_hash()
    if not force and SO_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == digest:
        return SO_PATH
    # several ranks of one torchrun job may find a stale library at the same time: serialise, then re-check
    import fcntl
    OBJ_DIR.mkdir(exist_ok=True)
    with open(OBJ_DIR / "build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and SO_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == digest:
                return SO_PATH
            return _build_locked(digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(digest: str, verbose: bool) -> Path:
    import torch
    from torch.utils import cpp_extension as ce

    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build the sm_100a extension")
    cuda_home = str(Path(nvcc).parent.parent)
    OBJ_DIR.mkdir(exist_ok=True)
    cu, cpp, _ = _sources()
    log = open(OBJ_DIR / "build.log", "w")

    def cc_cu(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        out = _run([nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-I", str(HERE), "-c", str(src), "-o", str(obj)], log)
        (OBJ_DIR / (src.stem + ".ptxas.txt")).write_text(out)
        return obj

    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"], "-I", str(HERE), "-isystem", cuda_home + "/include"]
    abi = int(torch.compiled_with_cxx11_abi())

    def cc_cpp(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
              "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, "-c", str(src), "-o", str(obj)], log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        futs = [ex.submit(cc_cpp, s) for s in cpp] + [ex.submit(cc_cu, s) for s in cu]
        objs = [f.result() for f in futs]

    torch_lib = str(Path(torch.__file__).parent / "lib")
    tmp = SO_PATH.with_suffix(f".so.tmp{os.getpid()}")
    _run(["g++", "-shared", "-o", str(tmp), *map(str, objs), f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}",
          "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
          f"-L{cuda_home}/lib64", f"-Wl,-rpath,{cuda_home}/lib64", "-lcudart"], log)
    os.replace(tmp, SO_PATH)
    HASH_PATH.write_text(digest)
    log.close()
    if verbose:
        print(f"[tds build] built {SO_PATH}")
    return SO_PATH


def _import_so():
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("tiny_deepspeed_b200._C", str(SO_PATH))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["tiny_deepspeed_b200._C"] = mod
    return mod


def load():
    """Import ``_C`` (rebuilding if stale and possible).  Raises if it cannot be provided."""
    if "tiny_deepspeed_b200._C" in sys.modules:
        return sys.modules["tiny_deepspeed_b200._C"]
    stale = not (SO_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == source_hash())
    if stale:
        if nvcc_path() is not None:
            build()
        elif not SO_PATH.exists():
            from ..ops._dispatch import ExtensionMissing
            raise ExtensionMissing(
                "tiny_deepspeed_b200/_C.so is missing and nvcc is unavailable; run "
                "`python -c 'import __graft_entry__ as g; g.build()'` where nvcc exists. "
                "CUDA tensors never fall back to PyTorch kernels.")
    return _import_so()
