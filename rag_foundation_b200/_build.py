"""Build recipe for librf_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build().

Each translation unit is compiled to an object under build/obj (in parallel, only when it or a header
is newer), then linked into the in-tree shared library that travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "librf_b200.so")
SOURCES = ["engine.cu", "group.cu", "score_topk.cu", "score_topk_gemm.cu", "score_topk_gemm_pair.cu", "featurize.cu", "synth.cu",
           "peak_probe.cu", "hostcopy.cpp"]
HEADERS = ["rf_device.cuh", "rf_gemm_device.cuh", "rf_internal.h", os.path.join("..", "..", "include", "rf_b200.h")]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [*ARCH_FLAGS, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: librf_b200.so cannot be built (there is no CPU fallback)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _env():
    env = dict(os.environ)
    # the image's CC/CXX wrappers lack some spec files; let nvcc find the distro g++
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources() + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile_one(src: str, force: bool, verbose: bool):
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps if os.path.exists(d)):
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True, env=_env())
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
    return obj, res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        results = list(pool.map(lambda s: _compile_one(s, force, verbose), srcs))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    cmd = [_nvcc(), *ARCH_FLAGS, "-shared", "-o", LIB, *[o for o, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True, env=_env())
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB


def sass_table() -> str:
    """`cuobjdump -sass` mnemonic counts per kernel of the built library: the reproducible evidence that the
    contraction and tile-movement kernels are Blackwell-native (UTCIMMA = tcgen05.mma, UTMALDG = TMA tensor
    load, UBLKCP = bulk copy, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit)."""
    import re
    cuobjdump = os.path.join(os.path.dirname(_nvcc()), "cuobjdump")
    out = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, env=_env()).stdout
    wanted = ["UTCIMMA", "UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "IDP.4A", "IDP", "REDUX", "SHFL", "LDGSTS", "ACQBULK"]
    rows, cur, counts = [], None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                rows.append((cur, counts))
            cur, counts = m.group(1), {}
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for w in wanted:
                if op == w or op.startswith(w + "."):
                    counts[w] = counts.get(w, 0) + 1
                    break
    if cur:
        rows.append((cur, counts))
    demangle = shutil.which("c++filt")
    lines = ["# cuobjdump -sass mnemonic counts per kernel of librf_b200.so (sm_100a)", ""]
    for name, c in sorted(rows):
        pretty = name
        if demangle:
            pretty = subprocess.run([demangle, name], capture_output=True, text=True).stdout.strip() or name
        pretty = re.sub(r"\(.*", "", pretty.replace("(anonymous namespace)::", "").replace("void ", ""))
        lines.append(f"{pretty}: " + (", ".join(f"{k} {v}" for k, v in sorted(c.items())) or "-"))
    return "\n".join(lines) + "\n"
