"""The C-ABI library loads without a GPU and exports every symbol include/frcnn_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frcnn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from tf_faster_rcnn_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header is not exported" % n
        assert n in _native.SIGNATURES, "symbol %s has no ctypes signature" % n
    for n in _native.SIGNATURES:
        assert n in names, "ctypes signature for undeclared symbol %s" % n


def test_version_and_error_text_without_gpu():
    from tf_faster_rcnn_b200 import _native
    L = _native.lib()
    assert L.frcnn_version() >= 100
    import torch
    if not torch.cuda.is_available():
        rc = L.frcnn_check_device(0)
        assert rc != 0 and "device" in _native.last_error().lower()      # fails loudly, no fallback
        # argument validation does not need a device
        rc = L.frcnn_conv_plan_create(None, None)
        assert rc == -2 and _native.last_error()


def test_conv_desc_struct_matches_header():
    """field count/order of the ctypes mirror == the C struct (guards silent ABI drift)."""
    from tf_faster_rcnn_b200 import _native
    src = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} frcnn_conv_desc;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(float\*|void\*|float|int)\s*", "", decl)
        fields += [f.strip().lstrip("*") for f in decl.split(",")]
    assert fields == [f[0] for f in _native.ConvDesc._fields_]


def test_every_entry_rejects_null_arguments_before_touching_a_device():
    """All-null / all-zero arguments: each status-returning entry answers ERR_ARG (-2) with a message -- it validates
    before any CUDA call, so this runs (and must not crash) on a machine without a GPU."""
    from tf_faster_rcnn_b200 import _native
    L = _native.lib()
    skip = {"frcnn_version", "frcnn_last_error", "frcnn_check_device", "frcnn_sort_workspace_bytes", "frcnn_detect_post_workspace_bytes",
            "frcnn_conv_plan_destroy", "frcnn_graph_destroy"}
    checked = 0
    for name in _native.SIGNATURES:
        if name in skip:
            continue
        fn = getattr(L, name)
        args = [None if t is ctypes.c_void_p else t() for t in fn.argtypes]
        rc = fn(*args)
        assert rc == -2, (name, rc)
        assert _native.last_error(), name
        checked += 1
    assert checked >= 18
    assert L.frcnn_sort_workspace_bytes(0) >= 0
    L.frcnn_conv_plan_destroy(None)                       # destroying nothing is a no-op
    L.frcnn_graph_destroy(None)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import it."""
    offenders = []
    for base in ("tf_faster_rcnn_b200", "tools"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(d, f)).read()
                    if re.search(r"^\s*(from\s+oracle\b|import\s+oracle\b)", src, flags=re.M):
                        offenders.append(os.path.join(d, f))
    assert offenders == []


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/frcnn_b200.h compiles as C99 (-pedantic, no warnings) and examples/nms_from_c.c -- the `_nms` call a C / Cython
    maintainer would write -- links against the library and runs: version + empty-input call succeed without
    a device; with a B200 present the 4-box call must keep boxes 0 and 2, without one it must fail with a message."""
    import subprocess
    exe = str(tmp_path / "nms_from_c")
    libdir = os.path.join(ROOT, "tf_faster_rcnn_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "nms_from_c.c"), "-L" + libdir, "-lfrcnn_b200", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "empty input: status 0, kept 0" in r.stdout
    assert ("kept 2: 0 2" in r.stdout) or ("status -" in r.stdout)
