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
    assert _lib.lib.glove_abi_version() == 1


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
