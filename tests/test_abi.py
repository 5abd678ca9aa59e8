"""The C-ABI library loads and exports every symbol include/mtbc.h declares (no compute calls: CPU only)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _decls():
    src = open(os.path.join(ROOT, "include", "mtbc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(?:int64_t|int|void|double|const char\*)\s+(mtbc_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)


def test_header_symbols_exported(lib):
    decls = _decls()
    assert len(decls) >= 45
    for name, _ in decls:
        assert hasattr(lib, name), f"{name} declared in mtbc.h but not exported by libmtbc.so"


def test_ctypes_signatures_match_header(lib):
    from multi_task_breast_cancer_b200 import _lib
    decls = dict(_decls())
    special = {"mtbc_last_error", "mtbc_build_digest"}
    assert set(_lib._SIGS) | special == set(decls), set(decls) ^ (set(_lib._SIGS) | special)
    for name, args in decls.items():
        if name in special:
            continue
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert n == len(_lib._SIGS[name]), (name, n, len(_lib._SIGS[name]))


def test_abi_version_and_error_string(lib):
    from multi_task_breast_cancer_b200 import _lib, build
    assert lib.mtbc_abi_version() == _lib.ABI_VERSION == 4
    assert isinstance(lib.mtbc_last_error(), bytes)
    # the digest compiled into the binary is the digest of the sources on disk (no side-car stamp file)
    assert build.lib_digest() == build._digest()


def test_stale_library_is_refused(tmp_path, monkeypatch):
    """A libmtbc.so built from other sources must not be bound silently (tracked stamp vs ignored .so)."""
    import pytest
    from multi_task_breast_cancer_b200 import _lib, build
    def no_compiler(*a, **k):
        raise RuntimeError("nvcc not found")
    monkeypatch.setattr(build, "_digest", lambda: "0" * 64)
    monkeypatch.setattr(build, "build", no_compiler)     # with a compiler the loader rebuilds instead (build.build)
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.MtbcError, match="other sources"):
        _lib.load()


def test_no_library_fallback_symbols():
    """The hot path may not reach cuDNN / cuBLAS: the shared object must not link them."""
    import subprocess
    from multi_task_breast_cancer_b200 import _lib
    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "cudnn" not in out and "cublas" not in out, out


def test_invalid_descriptor_is_an_error_not_a_crash(lib):
    import ctypes as C
    from multi_task_breast_cancer_b200 import _lib
    d = _lib.ConvGemmDesc()
    h = C.c_void_p()
    rc = lib.mtbc_conv_gemm_create(C.byref(d), C.byref(h))
    assert rc != 0 and b"conv_gemm" in lib.mtbc_last_error()


def test_product_package_never_imports_the_oracle_or_aten_compute():
    """The oracle is test infrastructure: nothing under multi_task_breast_cancer_b200/ may import it, and the package
    may not reach ATen's conv / norm / pooling kernels (the hot path is libmtbc.so or an error)."""
    import ast
    pkg = os.path.join(ROOT, "multi_task_breast_cancer_b200")
    banned_calls = {"conv2d", "conv_transpose2d", "instance_norm", "batch_norm", "max_pool2d", "leaky_relu",
                    "interpolate", "adaptive_avg_pool2d", "linear"}
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Import):
                assert not any(a.name.split(".")[0] == "oracle" for a in node.names), fn
            if isinstance(node, ast.ImportFrom):
                assert (node.module or "").split(".")[0] != "oracle", fn
                assert (node.module or "") != "torch.nn.functional", fn
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
                assert node.func.attr not in banned_calls, (fn, node.func.attr, node.lineno)


def test_descriptor_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the descriptor structs have the size and field offsets the C compiler gives include/mtbc.h
    (a field appended on one side only would shift every later argument silently)."""
    import ctypes as C
    import shutil
    import subprocess
    import pytest
    from multi_task_breast_cancer_b200 import _lib as L
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        pytest.skip("no C compiler")
    pairs = [("mtbc_act_view", L.ActView, ["ptr", "sN"]), ("mtbc_gemm_seg", L.GemmSeg, ["wtap"]),
             ("mtbc_out_slice", L.OutSlice, ["accumulate"]),
             ("mtbc_conv_gemm_desc", L.ConvGemmDesc, ["seg", "wpack", "outs", "dtype", "wpack_lo", "bwd_y", "bwd_slope"]),
             ("mtbc_wgrad_desc", L.WgradDesc, ["taps", "dw_acc", "dtype"]),
             ("mtbc_wgrad_multi_desc", L.WgradMultiDesc, ["k0", "dy", "splits"]),
             ("mtbc_convT_bwd_desc", L.ConvTBwdDesc, ["dy", "wd", "dw_acc", "dbias", "dx", "Cout"]),
             ("mtbc_param_job", L.ParamJob, ["i", "src", "dst1"])]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "mtbc.h"', "int main(void) {"]
    for cname, _, fields in pairs:
        lines.append(f'  printf("%zu", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({cname}, {f}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for (cname, ct, fields), line in zip(pairs, out):
        want = [int(v) for v in line.split()]
        have = [C.sizeof(ct)] + [getattr(ct, f).offset for f in fields]
        assert have == want, (cname, fields, have, want)
