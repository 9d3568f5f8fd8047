"""The C-ABI library loads and exports every symbol include/splat_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "splat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SPL_API\s+[^;(]*?\b(spl_\w+)\s*\(", text)))


def test_header_declares_what_python_binds(splat):
    from splatapult_b200 import api

    declared = _declared_symbols()
    assert len(declared) >= 35
    assert sorted(api.EXPORTED_SYMBOLS) == declared


def test_library_exports_every_declared_symbol(splat):
    lib = ctypes.CDLL(splat.library_path())
    for name in _declared_symbols():
        assert getattr(lib, name) is not None, name


def test_only_the_c_abi_is_exported(splat):
    """-fvisibility=hidden: nothing but spl_* (and the C++ drop-in classes) leaks out of the .so."""
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", splat.library_path()], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    c_names = [n for n in names if not n.startswith("_Z")]
    assert c_names and all(n.startswith("spl_") for n in c_names), c_names
    # C++ symbols: exactly the drop-in classes of include/splatapult_b200/*.h, none of the internals (namespace splat)
    dem = subprocess.run(["nm", "-DC", "--defined-only", splat.library_path()], capture_output=True, text=True).stdout
    cpp = [l.split(" T ", 1)[1] for l in dem.splitlines() if " T " in l and not l.split(" T ", 1)[1].startswith("spl_")]
    classes = {n.split("::")[0] for n in cpp}
    assert classes == {"GaussianCloud", "SplatRenderer", "CamerasConfig", "Ply", "SplatComm"}, classes
    assert not any(n.startswith("splat::") for n in cpp)


def test_version_and_error_channel(splat):
    L = splat.lib()
    assert L.spl_version() >= 100
    assert L.spl_device_count() >= 0
    assert L.spl_cloud_import_ply(None, b"x") < 0
    assert b"null" in L.spl_last_error()


def test_no_gpu_means_loud_failure(splat):
    """without a CUDA device the renderer must fail, never fall back (there is no CPU path in the product)."""
    if splat.device_count() > 0:
        pytest.skip("a GPU is present")
    gc = splat.GaussianCloud()
    gc.InitDebugCloud()
    r = splat.SplatRenderer()
    assert not r.Init(gc, False, False)
    assert "no CPU fallback" in r.last_error or "CUDA" in r.last_error or "cuda" in r.last_error
    with pytest.raises(splat.SplatError):
        r.Sort(splat.translate(0, 0, 3), splat.perspective(splat.FOVY, 1.0), [0, 0, 64, 64], [0.1, 1000.0])
    with pytest.raises(splat.SplatError):
        splat.sort_pairs_host([1, 2, 3], [0, 1, 2])


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under splatapult_b200/ may reference it."""
    pkg = os.path.join(ROOT, "splatapult_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("import ", "from ", "#include")):
                        assert "oracle" not in s, (f, s)


def test_header_is_plain_c_and_links(splat, scratch):
    """include/splat_b200.h compiles as C99 (-pedantic -Werror) and a C program links against the library."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(splat.library_path())
    exe = os.path.join(str(scratch), "c_abi_consumer")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "cpp", "c_abi_consumer.c"), "-L", libdir, "-lsplat_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = os.path.join(str(scratch), "pruned.ply")
    r = subprocess.run([exe, out], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("c abi ok") and os.path.getsize(out) > 0


def test_multigpu_entry_points_without_a_gpu(splat):
    """NCCL is bound at run time; the NUMA helpers degrade to 'unknown' -- and nothing falls back to a CPU path."""
    L = splat.lib()
    v = splat.SplatComm.NcclVersion()
    assert v == 0 or v >= 21800  # 0: no libnccl.so.2 on this machine; otherwise a real version (2.18+)
    if splat.device_count() > 0:
        pytest.skip("a GPU is present")
    assert splat.device_numa_node(0) == -1
    assert splat.bind_thread_near(0) == -1
    assert not L.spl_host_alloc_near(4096, 0)  # pinned memory needs the CUDA driver
    if v:
        uid = splat.SplatComm.GetUniqueId()
        assert len(uid) == 128 and any(uid)
        with pytest.raises(splat.SplatError):
            splat.SplatComm(uid, 0, 1, 0)  # ncclCommInitRank needs a device: loud failure
