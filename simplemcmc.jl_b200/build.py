"""Builds libsimplemcmc_cuda.so in-tree with nvcc for sm_100a (B200).  No JIT cache: the .so travels with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsimplemcmc_cuda.so")
SOURCES = ["smc_core.cu", "vm_exec.cu", "glm_fused.cu", "glm_i8.cu", "samplers.cu", "smc_comm.cu", "smc_p2p.cu", "affnorm.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default", "--use_fast_math=false"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in ("simplemcmc.h", "simplemcmc_tape.h")]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


TOOLS = ["i8_umma_probe", "fp64_issue_probe", "fp64_vs_imma_probe", "cond_graph_probe"]


def build_tools(force=False):
    """stand-alone sm_100a probes under tools/ (the tcgen05 int8 instruction probe, the FP64 issue probes and the CUDA-graph
    WHILE-node probe of DESIGN.md section 8); not linked into the library"""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tools = os.path.join(HERE, "..", "tools")
    out = []
    for name in TOOLS:
        src, exe = os.path.join(tools, name + ".cu"), os.path.join(tools, name)
        if os.path.exists(src) and (force or _stale(exe, [src])):
            subprocess.check_call([nvcc, "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src])
        out.append(exe)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
