"""Builds libusb200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m usearch12_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libusb200.so")
SOURCES = ["usb_api.cu", "usb_hostindex.cpp", "usb_udbfile.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "-Xptxas", "-v",
]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in os.listdir(root):
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".inc")):
                m = max(m, os.path.getmtime(os.path.join(root, fn)))
    return m


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing or stale" % LIB)
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


CLI = os.path.join(HERE, "usearch12_b200_cli")


def build_cli(force=False):
    """Host C++ driver (usearch12_b200/csrc/host) linked against libusb200.so."""
    host = os.path.join(CSRC, "host")
    srcs = [os.path.join(host, f) for f in ("usb_host.cpp", "usb_cluster_host.cpp", "usb_main.cpp")]
    newest = max(os.path.getmtime(x) for x in srcs + [os.path.join(host, "usb_host.h")])
    if not force and os.path.exists(CLI) and os.path.getmtime(CLI) >= newest:
        return CLI
    build()
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", CLI] + srcs + [
        "-L" + HERE, "-lusb200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building the host CLI")
    return CLI


REPLAY = os.path.join(HERE, "format_replay")


def build_format_replay(force=False):
    """tools/format_replay.cpp: test tool that feeds hit tables to the host sinks (no device needed)."""
    host = os.path.join(CSRC, "host")
    srcs = [os.path.join(os.path.dirname(HERE), "tools", "format_replay.cpp")] + [
        os.path.join(host, f) for f in ("usb_host.cpp", "usb_cluster_host.cpp")]
    newest = max(os.path.getmtime(x) for x in srcs + [os.path.join(host, "usb_host.h")])
    if not force and os.path.exists(REPLAY) and os.path.getmtime(REPLAY) >= newest:
        return REPLAY
    build()
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", REPLAY] + srcs + [
        "-L" + HERE, "-lusb200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building tools/format_replay.cpp")
    return REPLAY


if __name__ == "__main__":
    build_cli(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose=True))
