"""The C-ABI library loads without a GPU, exports every symbol include/tta.h declares, and the
numpy struct mirrors in tta_runtime.py have the C layouts (checked with a gcc-compiled probe)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

import tta_runtime as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'tta.h')


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tta_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(rt.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(rt.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(rt.EXPORTS)
    lib.tta_version.restype = ctypes.c_int
    assert lib.tta_version() >= 100


def test_struct_layouts_match_header(tmp_path):
    fields = {
        'tta_ew_task': rt.EW_TASK, 'tta_fold_task': rt.FOLD_TASK, 'tta_gram_task': rt.GRAM_TASK,
        'tta_eig_task': rt.EIG_TASK, 'tta_select_task': rt.SELECT_TASK, 'tta_gemm_task': rt.GEMM_TASK,
        'tta_sqnorm_task': rt.SQNORM_TASK, 'tta_refine_task': rt.REFINE_TASK}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tta.h"', 'int main(void){']
    for sname, dt in fields.items():
        lines.append('printf("%s %zu\\n", "{0}", sizeof({0}));'.format(sname))
        for f in dt.names:
            lines.append('printf("%s.%s %zu\\n", "{0}", "{1}", offsetof({0}, {1}));'.format(sname, f))
    lines += ['return 0;}']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split('\n')
    got = dict(l.split() for l in out if l.strip())
    for sname, dt in fields.items():
        assert int(got[sname]) == dt.itemsize, sname
        for f in dt.names:
            assert int(got[sname + '.' + f]) == dt.fields[f][1], (sname, f)


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(rt, 'LIB_PATH', '/nonexistent/libtta.so')
    monkeypatch.setattr(rt, '_LIB', None)
    with pytest.raises(rt.TtaError):
        rt.lib()


def test_cpu_tensors_are_rejected_without_emulator():
    import torch
    assert not rt.backend_is_emulated()
    with pytest.raises(rt.TtaError):
        rt.require_device(torch.zeros(4))
