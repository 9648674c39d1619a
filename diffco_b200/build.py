"""Build recipe for libdiffco_b200.so (hand-written sm_100a CUDA behind a C ABI; include/diffco_b200.h).

In-tree build with plain nvcc: ``python -m diffco_b200.build``.  The shared object lands next to the sources
(diffco_b200/csrc/libdiffco_b200.so) so that it travels to the GPU box with the repository snapshot.  The
thread-per-query kernel is compiled once per (radial kind, class width, mode) from dc_score_tq_inst.cu so the
instantiations build in parallel.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(CSRC, "libdiffco_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", CSRC] + ARCH_FLAGS

TQ_KINDS = {"rq2": "KR_RQ2", "ph1": "KR_PH1", "mq": "KR_MQ"}
TQ_VARIANTS = [("c1_score", 1, "M_SCORE"), ("c1_grad", 1, "M_GRAD"), ("c4_score", 4, "M_SCORE"), ("c4_grad", 4, "M_GRAD")]


def _units():
    units = [(n, os.path.join(CSRC, n + ".cu"), []) for n in ("dc_api", "dc_train", "dc_host", "dc_traj", "dc_peer", "dc_score_tc", "dc_score_ls_f32", "dc_score_ls_f64")]
    for kname, kenum in TQ_KINDS.items():
        for vname, cw, mode in TQ_VARIANTS:
            sym = f"tq_{kname}_{vname}"
            units.append((sym, os.path.join(CSRC, "dc_score_tq_inst.cu"),
                          [f"-DDC_TQ_KIND={kenum}", f"-DDC_TQ_CW={cw}", f"-DDC_TQ_MODE={mode}", f"-DDC_TQ_NAME={sym}"]))
    return units


def _sources_digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".txt")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(COMMON).encode())
    return h.hexdigest()


def _compile(unit, verbose):
    name, src, defs = unit
    obj = os.path.join(OBJ_DIR, name + ".o")
    cmd = [NVCC] + COMMON + defs + ["-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    """Compile every CUDA translation unit for sm_100a and link libdiffco_b200.so.  Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "sources.sha256")
    digest = _sources_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB_PATH}")
    units = _units()
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        results = list(ex.map(lambda u: _compile(u, verbose), units))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    link = [NVCC, "-shared", "-o", LIB_PATH] + objs + ARCH_FLAGS + ["-cudart", "static", "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
