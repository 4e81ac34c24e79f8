"""CPU: the C-ABI library loads and exports every symbol include/tapenv.h declares; host-side
configuration logic (no kernels are launched here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tapenv.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tapenv_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tapenv import _capi
    names = declared_symbols()
    assert len(names) >= 14
    raw = C.CDLL(_capi._build.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libtapenv.so does not export %s" % n
        assert n in _capi.SYMBOLS, "%s has no ctypes prototype" % n
    assert _capi.lib.tapenv_version() == 100 + 0 or _capi.lib.tapenv_version() > 100


def test_config_from_reference_option_strings():
    from tapenv import make_config, _capi
    c = make_config(4096, 10, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", "bot", True)
    assert (c.dim, c.rotate_types, c.width, c.length, c.height) == (2, 2, 5, 1, 50)
    assert (c.static_rows, c.dyn_rows, c.update_time) == (3, 30, 3)
    assert c.reward_flags == 2 | 4 and c.ratio_mode == 4 and c.strategy == _capi.LB_GREEDY
    c = make_config(8, 10, [5, 5, 50], "C+P+S-lb-hard", "full", "LB_GREEDY", "bot", True)
    assert (c.dim, c.rotate_types, c.length, c.static_rows, c.dyn_rows) == (3, 6, 5, 4, 30)
    assert c.reward_flags & 1
    # tools.py:3617-3620: the reward type overrides the packing strategy
    c = make_config(8, 20, [7, 50], "C+P+S-mcs-hard", "diff", "LB_GREEDY", "bot", True)
    assert c.strategy == _capi.MACS and c.reward_flags & 8 and not c.reward_flags & 16
    # tools.py:3961: C+P-lb-soft -> (C+P)/2 ; no 'S' in the string -> S term off
    c = make_config(8, 10, [5, 50], "C+P-lb-soft", "zero", "LB_GREEDY", "bot", True)
    assert c.ratio_mode == 7 and not c.reward_flags & 4
    c = make_config(8, 10, [5, 5, 50], "C+P+S-lb-soft", "diff", "LB", "bot", True)
    assert c.strategy == _capi.LB
    lay = _capi.StateLayout()
    assert _capi.lib.tapenv_state_get_layout(C.byref(c), C.byref(lay)) == 0
    assert lay.lists - lay.voxels >= 8 * 25 * 50 * 2 and lay.pending - lay.lists >= 8 * 50 * 5 * 12
    c = make_config(8, 10, [5, 5, 50], "C+P+S-mcs-hard", "diff", "LB_GREEDY", "bot", True)      # reward type forces MACS, 3D: voxel state
    assert c.strategy == _capi.MACS
    assert _capi.lib.tapenv_state_get_layout(C.byref(c), C.byref(lay)) == 0
    assert lay.lists - lay.voxels >= 8 * 25 * 50 * 2 and lay.pending - lay.lists >= 8 * 50 * 5 * 12
    c = make_config(8, 10, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", "simple", False)
    assert (c.rotate_types, c.dyn_rows, c.update_time) == (1, 10, 1)
    c = make_config(8, 10, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", "rot-old", True)         # legacy: + one rotate-state row
    assert (c.static_rows, c.dyn_rows, c.update_time, c.rotate_types) == (3, 11, 1, 2)
    c = make_config(8, 10, [5, 5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", "mul-with", True)     # + target-container row
    assert (c.static_rows, c.dyn_rows, c.update_time, c.rotate_types) == (5, 30, 3, 6)


@pytest.mark.parametrize("kw,code", [
    (dict(reward_type="bogus"), -2), (dict(heightmap_type="bogus"), -2), (dict(packing_strategy="bogus"), -2),
    (dict(input_type="bogus"), -2), (dict(container_size=[64, 50]), -3), (dict(blocks_num=40), -3),
    (dict(container_size=[5, 5, 300], packing_strategy="MACS"), -3),
])
def test_config_errors(kw, code):
    from tapenv import make_config, TapEnvError
    args = dict(batch=4, blocks_num=10, container_size=[5, 50], reward_type="C+P+S-lb-soft", heightmap_type="diff",
                packing_strategy="LB_GREEDY", input_type="bot", allow_rot=True)
    args.update(kw)
    with pytest.raises(TapEnvError) as e:
        make_config(**args)
    assert e.value.code == code and isinstance(e.value, ValueError)


def test_state_layout_is_disjoint_and_sized():
    from tapenv import make_config, _capi
    c = make_config(1000, 10, [5, 5, 50], "C+P+S-lb-soft", "diff")
    lay = _capi.StateLayout()
    assert _capi.lib.tapenv_state_get_layout(C.byref(c), C.byref(lay)) == 0
    offs = [lay.scalars, lay.heightmap, lay.positions, lay.blocks, lay.stable, lay.flags, lay.voxels, lay.lists, lay.pending, lay.total]
    assert lay.voxels == lay.lists == lay.pending == lay.total       # LB-only sections are empty for LB_GREEDY
    assert offs == sorted(offs) and lay.total == _capi.lib.tapenv_state_bytes(C.byref(c))
    assert lay.heightmap - lay.scalars >= 1000 * 16 and lay.positions - lay.heightmap >= 1000 * 25 * 4
    assert _capi.lib.tapenv_encoded_heightmap_len(C.byref(c)) == 50
    assert all(o % 16 == 0 for o in offs)


def test_null_and_cpu_arguments_are_rejected_without_launch():
    import torch
    from tapenv import make_config, _capi, update_dynamic
    c = make_config(4, 10, [5, 50])
    assert _capi.lib.tapenv_reset(C.byref(c), None, None, None, None, None) == _capi.EINVAL
    assert _capi.lib.tapenv_step(C.byref(c), None, None, None, None, None, None, None, None, None, None, None) == _capi.EINVAL
    with pytest.raises(RuntimeError):                       # no CPU fallback
        update_dynamic(torch.zeros(4, 30, 20), torch.zeros(4, 3, 20), torch.zeros(4, dtype=torch.long), "bot", True)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under tap-net_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "tap-net_b200")
    bad = re.compile(r"(from|import)\s+oracle|tap_oracle|libtap_oracle|oracle[/.](oracle|refshim|_build|_ref)|refshim")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f)).read()
                assert not bad.search(src), os.path.join(d, f)


def test_pack_inputs_layout():
    """Host-side compaction used by the packed upload path: u8 static, little-endian bit rows padded to 16 bytes."""
    import numpy as np
    from tapenv.dataset import pack_inputs, packed_words
    from tests.golden_io import load_inputs
    static, dynamic = load_inputs("rand3d_n10.npz", 5)
    su8, bits = pack_inputs(static, dynamic)
    B, rows, S = dynamic.shape
    assert su8.dtype == np.uint8 and np.array_equal(su8.astype(np.float32), static)
    assert bits.dtype == np.int32 and bits.shape == (B, packed_words(rows, S)) and bits.shape[1] % 4 == 0
    flat = dynamic.reshape(B, -1)
    for q in (0, 1, 31, 32, 777, rows * S - 1):
        assert np.array_equal((bits[:, q >> 5] >> (q & 31)) & 1, flat[:, q].astype(np.int32))
    u = bits.view(np.uint32)
    assert int(sum(bin(int(v)).count("1") for v in u.reshape(-1))) == int(flat.sum())      # padding bits are zero
    import tapenv
    cfg = tapenv.make_config(B, 10, [5, 5, 50])
    assert int(tapenv._capi.lib.tapenv_packed_words(C.byref(cfg))) == bits.shape[1]


def test_host_batch_layout_is_aligned_and_disjoint():
    """The staging layout shared by HostPipeline's device buffers and HostBatch (one contiguous upload): 256-byte aligned
    sections, disjoint, exact views."""
    import torch
    from tapenv.runner import _carve, _carve_bytes
    specs = [(torch.float32, (1, 7, 3, 20)), (torch.float32, (1, 7, 30, 20)), (torch.int64, (1, 10, 7)), (torch.uint8, (7, 3, 20))]
    nbytes = _carve_bytes(specs)
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    views = _carve(buf, specs)
    base = buf.data_ptr()
    spans = []
    for v, (dt, shape) in zip(views, specs):
        assert v.dtype == dt and tuple(v.shape) == tuple(shape) and v.is_contiguous()
        off = v.data_ptr() - base
        assert off % 256 == 0 and off + v.numel() * v.element_size() <= nbytes
        spans.append((off, off + v.numel() * v.element_size()))
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    views[2].fill_(5)
    assert int(buf.sum()) == 5 * 70 and int(views[0].abs().sum()) == 0      # writes land only in their own section
