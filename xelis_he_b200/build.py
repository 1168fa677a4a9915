"""Build libxhe_cuda.so (sm_100a only) in-tree with nvcc.  Called by __graft_entry__.build().

The library is rebuilt whenever the SHA-256 of its sources (csrc/, host/, include/ and the compiler flags) differs from the
hash recorded beside the last build (build/libxhe_cuda.srchash): a prebuilt .so that travels with the tree is only reused
if it was built from exactly these sources, whatever the file times say."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libxhe_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3,-pthread,-march=x86-64-v3",
         "-Xptxas", "-v", "--threads", "0"]


def sources():
    host = os.path.join(HERE, "host")
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu")) + sorted(os.path.join(host, f) for f in os.listdir(host) if f.endswith(".cpp"))


def _dep_files():
    return sorted(os.path.join(r, f) for d in (CSRC, os.path.join(HERE, "host"), os.path.join(HERE, "..", "include"))
                  for r, _, fs in os.walk(d) for f in fs if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp")) and "experimental" not in r)


def source_hash():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in _dep_files():
        h.update(os.path.relpath(p, HERE).encode()); h.update(open(p, "rb").read())
    return h.hexdigest()


def _includes_of(path, seen=None):
    """the file and every local header it includes, transitively (#include "...")"""
    import re
    seen = seen if seen is not None else set()
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for m in re.finditer(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), re.M):
        _includes_of(os.path.join(os.path.dirname(path), m.group(1)), seen)
    return seen


HASH_FILE = os.path.join(HERE, "build", "libxhe_cuda.srchash")


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(HASH_FILE):
        return True
    return open(HASH_FILE).read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    want = source_hash()
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(p) for p in _includes_of(src)):
            continue
        log = open(obj + ".log", "w")
        procs.append((src, subprocess.Popen([NVCC, *FLAGS, "-c", src, "-o", obj], stdout=log, stderr=subprocess.STDOUT), log))
    for src, p, log in procs:
        rc = p.wait()
        log.close()
        if rc != 0 or verbose:
            sys.stderr.write(open(log.name).read())
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart", "-Xcompiler", "-pthread"])
    with open(HASH_FILE, "w") as f:
        f.write(want + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
