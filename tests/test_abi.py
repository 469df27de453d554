"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/glove_b200.h
declares, the ctypes mirror of the argument struct matches the C layout, and argument errors are reported without a GPU."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "glove_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glove_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from glove_tensorflow_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(_lib.lib, n), "libglove_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "_lib.SIGNATURES has no prototype for %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib.glove_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layouts_match_the_c_header():
    from glove_tensorflow_b200 import _lib
    fields = [n for n, _ in _lib.StepArgs._fields_]
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"glove_b200.h\"\nint main(){printf(\"%zu %zu\", sizeof(glove_step_args), sizeof(glove_scalars));" + \
           "".join('printf(" %%zu", offsetof(glove_step_args, %s));' % f for f in fields) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(td, "t")])
        out = subprocess.check_output([os.path.join(td, "t")]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.StepArgs)
    assert int(out[1]) == ctypes.sizeof(_lib.GloveScalars) == 32
    for f, off in zip(fields, out[2:]):
        assert getattr(_lib.StepArgs, f).offset == int(off), f


def test_integration_md_binding_stub_matches_the_header():
    """The ctypes stub a maintainer would paste from INTEGRATION.md must describe the struct the library reads: same
    fields in the same order, same size and offsets as the compiled header (round 1 shipped a stub 16 bytes short)."""
    from glove_tensorflow_b200 import _lib
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class StepArgs\(ctypes\.Structure\):.*?\n(    _fields_ = \[.*?\])\n", md, flags=re.S)
    assert m, "INTEGRATION.md lost its StepArgs stub"
    ns = dict(ctypes=ctypes, vp=ctypes.c_void_p, i32=ctypes.c_int32, i64=ctypes.c_int64, f32=ctypes.c_float,
              u32=ctypes.c_uint32, sz=ctypes.c_size_t)
    exec("class StepArgs(ctypes.Structure):\n" + m.group(1), ns)
    Stub = ns["StepArgs"]
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct glove_step_args \{(.*?)\} glove_step_args;", src, flags=re.S).group(1)
    c_fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            c_fields += [n.strip().lstrip("*") for n in re.sub(r"^(const\s+)?\w+\s+", "", decl).split(",")]
    assert [n for n, _ in Stub._fields_] == c_fields == [n for n, _ in _lib.StepArgs._fields_]
    assert ctypes.sizeof(Stub) == ctypes.sizeof(_lib.StepArgs) == _lib.lib.glove_step_args_size()
    for n, _ in Stub._fields_:
        assert getattr(Stub, n).offset == getattr(_lib.StepArgs, n).offset, n   # _lib.StepArgs is checked against gcc above


def test_stale_step_args_are_refused():
    """struct_size guards every glove_step_args entry point: a binding compiled against another layout gets GLOVE_EINVAL
    with a message, not a read past its struct."""
    from glove_tensorflow_b200 import _lib
    args = _lib.StepArgs()
    args.struct_size = ctypes.sizeof(_lib.StepArgs) - 16          # round 1's stub
    args.row_table = args.col_table = args.scalars = args.plan = args.workspace = 8   # non-null, never dereferenced
    args.V, args.d, args.B, args.plan_K = 10, 4, 8, 1
    args.optimizer = 2
    assert _lib.lib.glove_train_step(ctypes.byref(args), None) == _lib.EINVAL
    assert b"struct_size" in _lib.lib.glove_last_error()


def test_size_queries_and_layout_helpers():
    from glove_tensorflow_b200._lib import lib, OPTIMIZERS
    assert lib.glove_table_stride(300) == 304 and lib.glove_table_stride(64) == 72 and lib.glove_table_stride(6) == 8
    assert [lib.glove_table_planes(OPTIMIZERS[o]) for o in ("Adam", "Adagrad", "SGD")] == [3, 2, 1]
    assert lib.glove_plan_bytes(0, 10) == 0 and lib.glove_step_workspace_bytes(0, 8) == 0
    a, b = lib.glove_plan_bytes(4, 1024), lib.glove_plan_bytes(8, 1024)
    assert 0 < a < b
    assert lib.glove_prepare_workspace_bytes(4, 1024) > 4 * 1024 * 4 * 10
    assert lib.glove_step_workspace_bytes(65536, 300) >= 2 * 65536 * 304 * 4
    assert lib.glove_eval_workspace_bytes(1000, 100) > 0 and lib.glove_host_staging_bytes(2, 128) > 2 * 128 * 24


def test_argument_errors_are_reported_without_a_gpu():
    from glove_tensorflow_b200 import _lib
    rc = _lib.lib.glove_table_init(None, 10, 4, 0, 0, None)
    assert rc == _lib.EINVAL and b"glove_table_init" in _lib.lib.glove_last_error()
    rc = _lib.lib.glove_prepare_batches(None, None, 0, None, None, None, None, 1, None, 0, 0, 0, 1, 1, 1, None)
    assert rc == _lib.EINVAL
    args = _lib.StepArgs()
    assert _lib.lib.glove_train_step(ctypes.byref(args), None) == _lib.EINVAL
    # shared plan construction: null plans, a world of one, a shard outside the world
    pull = _lib.lib.glove_plan_pull_slice
    assert pull(None, None, 4, 64, 2, 0, None) == _lib.EINVAL and b"glove_plan_pull_slice" in _lib.lib.glove_last_error()
    one = ctypes.c_void_p(256)
    assert pull(one, one, 4, 64, 1, 0, None) == _lib.EINVAL and pull(one, one, 4, 64, 4, 4, None) == _lib.EINVAL
    assert pull(one, one, 0, 64, 2, 0, None) == _lib.EINVAL and pull(one, one, 4, 64, 9, 0, None) == _lib.EINVAL
    with pytest.raises(_lib.GloveError):
        _lib.check(rc, "probe")


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from glove_tensorflow_b200.engine import GloveEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GloveEngine(10, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "glove_tensorflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libglove_oracle" not in src, f


def test_csv_schema_layout_matches_the_c_header():
    from glove_tensorflow_b200 import _lib
    prog = ('#include <stdio.h>\n#include <stddef.h>\n#include "glove_b200.h"\nint main(){printf("%zu %zu %zu %zu %d %d %d", '
            'sizeof(glove_csv_schema), offsetof(glove_csv_schema, n_cols), offsetof(glove_csv_schema, column), '
            'offsetof(glove_csv_schema, kind), GLOVE_CSV_TOKEN, GLOVE_CSV_INT, GLOVE_CSV_FLOAT);return 0;}')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(td, "t")])
        out = [int(x) for x in subprocess.check_output([os.path.join(td, "t")]).decode().split()]
    S = _lib.CsvSchema
    assert out[:4] == [ctypes.sizeof(S), S.n_cols.offset, S.column.offset, S.kind.offset]
    assert out[4:] == [_lib.CSV_TOKEN, _lib.CSV_INT, _lib.CSV_FLOAT]
