"""CPU checks of the C ABI: the library builds/loads without a GPU and exports every symbol include/gnca.h
declares; the ctypes structures mirror the header; argument validation never reaches a kernel."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from graph_neural_cellular_automata_b200 import _lib, build
    build.build(verbose=False)
    return _lib.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "gnca.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnca_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from graph_neural_cellular_automata_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libgnca.so does not export {n}"
        assert n in _lib.EXPORTS, f"_lib.py does not bind {n}"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_layout(lib):
    from graph_neural_cellular_automata_b200 import _lib
    from graph_neural_cellular_automata_b200 import functional as GF
    assert lib.gnca_version() == _lib.GNCA_VERSION
    hdr = open(os.path.join(ROOT, "include", "gnca.h")).read()
    assert f"#define GNCA_VERSION {_lib.GNCA_VERSION}" in hdr
    assert f"#define GNCA_MAX_K {_lib.GNCA_MAX_K}" in hdr
    d = GF.make_model_desc(16, 128, 16, graph=True, torus=True, hidden_only=True, alive_to_alive=True, groupnorm=True,
                           update_gain=0.05, alpha_thr=0.12, graph_alpha_thr=0.12)
    lay = GF.param_layout(d)
    # 9,169 floats receive a gradient in the graph model: 10,753 trainable - 1,584 dead gate_mlp
    assert lay.total == 9169
    assert (lay.w1, lay.b1, lay.w2, lay.gamma, lay.beta) == (0, 6144, 6272, 8320, 8336)
    dc = GF.make_model_desc(16, 128, 0, graph=False, torus=False, hidden_only=False, alive_to_alive=False,
                            groupnorm=True, update_gain=0.1, alpha_thr=0.1, graph_alpha_thr=0.1)
    assert GF.param_layout(dc).total == 8352 and GF.param_layout(dc).wm == -1
    assert GF.segment_offsets(d)[-1] == 9169 and len(GF.segment_offsets(d)) == 13


def test_argument_errors_do_not_launch(lib):
    from graph_neural_cellular_automata_b200 import _lib
    from graph_neural_cellular_automata_b200 import functional as GF
    d = GF.make_model_desc(5, 128, 16, graph=True, torus=True, hidden_only=True, alive_to_alive=True, groupnorm=True,
                           update_gain=0.05, alpha_thr=0.12, graph_alpha_thr=0.12)      # C=5: no kernel
    lay = _lib.GncaLayout()
    assert lib.gnca_param_layout(C.byref(d), C.byref(lay)) == -2
    assert b"unsupported" in lib.gnca_error_string(-2)
    assert lib.gnca_perception_fwd(1, 16, 8, 8, None, None, None) == -1
    assert lib.gnca_apply_mask(0, None, None, None) == -1
    before = lib.gnca_launch_count()
    assert lib.gnca_alive_mask(1, 3, 8, 8, C.c_void_p(16), 0.1, C.c_void_p(16), None) == -1     # C < 4
    assert lib.gnca_launch_count() == before


def test_struct_sizes_match_header():
    from graph_neural_cellular_automata_b200 import _lib
    assert C.sizeof(_lib.GncaModel) == 32
    assert C.sizeof(_lib.GncaLayout) == 14 * 8
    # gnca_schedule: 2 int32, 5 pointers, 2 uint64, 1 pointer, 2 int32
    assert C.sizeof(_lib.GncaSchedule) == 8 + 5 * 8 + 16 + 8 + 8 + 8
    assert C.sizeof(_lib.GncaDamage) == 24 + 16
