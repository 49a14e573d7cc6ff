#!/usr/bin/env python
"""Build the UNMODIFIED reference hot path into oracle/_ref/ (git-ignored, travels with gpurun).

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ may be imported by the product package
(vppstereo_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker or the CPU baseline.

Sources are compiled from where they lie under /root/reference; no reference source is
copied into the repo. Outputs (oracle/_ref/):
  pyrSGM.<abi>.so        <- RSGM/{pyrSGM,FastFilters,StereoBMHelper}.cpp   (g++ directly, flags of RSGM/setup.py:5
                            except -march=native -> -march=x86-64-v3 so the .so runs on the GPU box's host CPU)
  vpp_core_opt.<abi>.so  <- vpp_core/vpp_core_opt.pyx (cython -> C in a temp dir -> gcc -O2, baseline x86-64:
                            no FMA contraction, matching the reference's own default distutils build)
  rsgm_ref.pycode, vpp_standalone_ref.pycode, filter_ref.pycode
                         <- byte-compiled models/rsgm/rsgm.py, vpp_standalone.py, filter.py (glue + numba kernels)
  refmodels/{raft_stereo,psmnet}/*.pycode, losses_ref.pycode
                         <- byte-compiled models/raft_stereo, models/psmnet (the CONSUMERS of the projected images, BASELINE
                            configs[3]: random-initialised in the tests / bench, never rebuilt here) and losses.py (sample_hints)
Run every process that calls pyrSGM with MALLOC_MMAP_THRESHOLD_=65536: the reference
reads uninitialised malloc memory in census rows 0,1,H-2,H-1.
"""
import os, subprocess, sys, sysconfig, tempfile, py_compile, shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VPP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
RSGM = os.path.join(REF, "thirdparty/stereo-vision/reconstruction/base/rSGM")
CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False):
    if not os.path.isdir(REF):
        print(f"[oracle/build_ref] {REF} absent: keeping prebuilt oracle/_ref as is")
        return False
    import numpy
    os.makedirs(OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    gomp = ["-B/usr/lib/gcc/x86_64-linux-gnu/13"] if os.path.isdir("/usr/lib/gcc/x86_64-linux-gnu/13") else []

    so = os.path.join(OUT, "pyrSGM" + suffix)
    if force or not os.path.exists(so):
        _run([CXX, "-shared", "-fPIC", "-O3", "-ffast-math", "-msse4.1", "-msse4.2", "-march=x86-64-v3",
              "-fopenmp", "-Wno-write-strings", "-DNDEBUG", "-include", os.path.join(HERE, "numpy2_compat.h"),
              "-I" + RSGM] + inc + gomp +
             [os.path.join(RSGM, f) for f in ("pyrSGM.cpp", "FastFilters.cpp", "StereoBMHelper.cpp")] +
             ["-o", so])

    so = os.path.join(OUT, "vpp_core_opt" + suffix)
    if force or not os.path.exists(so):
        with tempfile.TemporaryDirectory() as tmp:
            c_file = os.path.join(tmp, "vpp_core_opt.c")
            _run([sys.executable, "-m", "cython", "-3", os.path.join(REF, "vpp_core/vpp_core_opt.pyx"), "-o", c_file])
            _run([CC, "-shared", "-fPIC", "-O2", "-fno-strict-overflow", "-DNDEBUG", "-w"] + inc + [c_file, "-o", so])

    for src, dst in (("models/rsgm/rsgm.py", "rsgm_ref.pycode"), ("vpp_standalone.py", "vpp_standalone_ref.pycode"),
                     ("filter.py", "filter_ref.pycode")):
        d = os.path.join(OUT, dst)
        if force or not os.path.exists(d):
            py_compile.compile(os.path.join(REF, src), cfile=d, doraise=True)
            print("+ py_compile", src, "->", d)
    d = os.path.join(OUT, "losses_ref.pycode")
    if (force or not os.path.exists(d)) and os.path.exists(os.path.join(REF, "losses.py")):
        py_compile.compile(os.path.join(REF, "losses.py"), cfile=d, doraise=True)
    # the two networks as sourceless packages (relative imports keep working): refmodels/<pkg>/<module>.pycode (imported through oracle/ref.py's finder)
    pk = os.path.join(OUT, "refmodels")
    for sub in ("raft_stereo", "raft_stereo/utils", "psmnet"):
        srcdir = os.path.join(REF, "models", sub)
        dstdir = os.path.join(pk, sub)
        os.makedirs(dstdir, exist_ok=True)
        names = [f for f in os.listdir(srcdir) if f.endswith(".py")]
        if "__init__.py" not in names:
            names.append("__init__.py")
        for f in names:
            d = os.path.join(dstdir, f[:-3] + ".pycode")     # (.pyc files do not travel with gpurun)
            if force or not os.path.exists(d):
                src = os.path.join(srcdir, f)
                if not os.path.exists(src):             # a directory without __init__.py (raft_stereo/utils): empty package marker
                    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as t:
                        src = t.name
                py_compile.compile(src, cfile=d, doraise=True)
    d = os.path.join(pk, "__init__.pycode")
    if force or not os.path.exists(d):
        with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as t:
            pass
        py_compile.compile(t.name, cfile=d, doraise=True)
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
