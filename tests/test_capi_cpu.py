"""CPU: the C-ABI library builds, loads and exports every symbol include/lvae_b200.h declares
(no compute calls without a GPU), and the package mirrors the reference's state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lvae_build", os.path.join(ROOT, "ladder-vae-pytorch_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lvae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lvae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    lib.lvae_abi_version.restype = ctypes.c_int
    assert lib.lvae_abi_version() == 1


def test_binding_covers_header(libpath):
    import lvae_b200
    assert sorted(lvae_b200._capi.exported_symbols()) == header_symbols()
    lvae_b200._capi.lib()       # loads and resolves every bound symbol


def header_prototypes():
    """name -> list of parameter kinds ('P' pointer, 'I' int, 'L' long long, 'U' unsigned long long, 'F' float) parsed from the header."""
    text = open(os.path.join(ROOT, "include", "lvae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"\b(lvae_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        kinds = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a or "lvae_stream_t" in a:
                    kinds.append("P")
                elif "unsigned long long" in a:
                    kinds.append("U")
                elif "long long" in a:
                    kinds.append("L")
                elif "float" in a:
                    kinds.append("F")
                elif "int" in a:
                    kinds.append("I")
                else:
                    raise AssertionError("unparsed parameter %r of %s" % (a, name))
        protos[name] = kinds
    return protos


def test_binding_signatures_match_header(libpath):
    """Every ctypes argtypes list has the arity and the parameter kinds of the prototype in include/lvae_b200.h (a pointer passed
    where the library expects a long long is silent corruption, not an exception)."""
    import lvae_b200
    capi = lvae_b200._capi
    kind = {ctypes.c_void_p: "P", ctypes.c_char_p: "P", ctypes.c_int: "I", ctypes.c_longlong: "L", ctypes.c_ulonglong: "U",
            ctypes.c_float: "F"}
    protos = header_prototypes()
    bound = dict(capi._SIGNATURES)
    bound.update({k: v[0] for k, v in capi._SPECIAL.items()})
    assert sorted(bound) == sorted(protos)
    for name, argtypes in bound.items():
        got = [kind[t] for t in argtypes]
        assert got == protos[name], "%s: binding %s vs header %s" % (name, "".join(got), "".join(protos[name]))


def test_no_cpu_fallback(libpath):
    import lvae_b200
    from oracle import lvae_oracle as O
    cfg = O.LVAEConfig(color_ch=1, z_dims=[4], img_shape=(8, 8), blocks_per_layer=1, n_filters=8, dropout=0.0,
                       likelihood_form="bernoulli", res_block_type="bacdbacd", merge_type="residual", gated=True)
    m = lvae_b200.LadderVAE(**cfg.kwargs())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(2, 1, 8, 8))


@pytest.mark.parametrize("name", ["mnist3", "cifar15"])
def test_state_dict_layout_matches_reference(name, libpath):
    import lvae_b200
    from oracle import lvae_oracle as O
    cfg = O.baseline_config(name)
    m = lvae_b200.LadderVAE(**cfg.kwargs())
    sd, shapes = m.state_dict(), O.param_shapes(cfg)
    assert list(sd.keys()) == list(shapes.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(shapes[k]), k
    m.load_state_dict(O.make_params(cfg, 1), strict=True)


def test_error_conventions(libpath):
    import lvae_b200
    from lvae_b200.lib.nn import ResidualBlock, ELU
    with pytest.raises(ValueError):
        ResidualBlock(8, ELU, block_type="nope")
    with pytest.raises(TypeError):
        ResidualBlock(8, ELU, block_type="bacdbacd", dropout=None)
    with pytest.raises(RuntimeError, match="Unrecognized likelihood"):
        lvae_b200.LadderVAE(1, [4], img_shape=(8, 8), likelihood_form="x", res_block_type="bacdbac", merge_type="residual")


def test_conv_spec_routing_and_pack_policy(libpath):
    """Host-side routing tables (no device work): which convolutions are eligible for the tcgen05 kernels and which packed
    weight layouts the engine refreshes per step."""
    import lvae_b200
    from lvae_b200.lib.nn import Conv2d, ConvTranspose2d
    res = Conv2d(64, 64, 3, padding=1).spec                       # residual conv
    gate = Conv2d(64, 128, 1).spec                                # gate conv
    merge = Conv2d(128, 64, 1).spec                               # two-input merge
    zout = Conv2d(32, 64, 3, padding=1).spec                      # conv_out of the stochastic block (input padded to 64)
    stem = Conv2d(3, 64, 5, padding=2, stride=2).spec
    down = Conv2d(64, 64, 3, padding=1, stride=2).spec
    up = ConvTranspose2d(64, 64, 3, padding=1, stride=2, output_padding=1).spec
    up_nopad = ConvTranspose2d(64, 64, 3, padding=1, stride=2).spec
    assert res.tc_shape and gate.tc_shape and merge.tc_shape and zout.tc_shape
    assert not stem.tc_shape and not down.tc_shape and not up.tc_shape
    assert down.s2_shape and up.s2_shape and not up_nopad.s2_shape and not res.s2_shape and not stem.s2_shape
    # bf16 pipeline: tensor-core convs only refresh their tcgen05 tiles; padded-input and CUDA-core convs keep the generic rows
    assert [p.mode for p in res.packs(True)] == [2, 3]
    assert [p.mode for p in gate.packs(True)] == [2, 3]
    assert sorted(p.mode for p in zout.packs(True)) == [0, 1, 2, 3]
    assert sorted(p.mode for p in stem.packs(True)) == [0, 1]
    assert sorted(p.mode for p in down.packs(True)) == [0, 1, 2, 3]
    assert sorted(p.mode for p in res.packs(False)) == [0, 1]
    # tcgen05 tile geometry: [tap][k-block][Npad][64]
    assert res.pack_tc_fwd.rows == 9 * 1 * 64 and gate.pack_tc_fwd.rows == 1 * 1 * 128 and merge.pack_tc_fwd.rows == 1 * 2 * 64
    # the two orientations of a stride-2 conv: Conv2d (co,ci) forward = gather (N=co), ConvTranspose2d (ci,co) forward = scatter (N=co)
    assert (down.pack_s2_gather.mode, down.pack_s2_scatter.mode) == (2, 3)
    assert (up.pack_s2_gather.O, up.pack_s2_gather.I) == (64, 64)
    assert down.out_hw(16, 16) == (8, 8) and up.out_hw(8, 8) == (16, 16)


def test_product_config_table_matches_oracle(libpath):
    """bench.py builds its models from lvae_b200.configs (the product never imports the oracle); the table must stay in
    step with the oracle's, which is the one pinned against the reference."""
    import lvae_b200
    from lvae_b200.configs import baseline_config, baseline_kwargs
    from oracle import lvae_oracle as O
    for name in ("mnist3", "mnist12", "cifar15", "celeba20"):
        ours, ref = baseline_kwargs(name), O.baseline_config(name).kwargs()
        assert set(ours) == set(ref)
        for k in ref:
            a, b = ours[k], ref[k]
            assert (list(a) == list(b)) if isinstance(b, (list, tuple)) else (a == b), (name, k, a, b)
        cfg = baseline_config(name)
        assert cfg.color_ch == ref["color_ch"] and tuple(cfg.img_shape) == tuple(ref["img_shape"]) and cfg.kwargs() == ours
    model = lvae_b200.LadderVAE(**baseline_kwargs("mnist3"))          # construction needs no device
    assert len(model.state_dict()) == len(O.param_shapes(O.baseline_config("mnist3"))) or len(model.state_dict()) > 100
