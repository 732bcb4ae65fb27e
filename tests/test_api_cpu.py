"""Host-side contract of the module mirror, checked without a device: sizes, shapes, attribute values,
parameter counts and exception types of lvae_b200.LadderVAE against the numbers the survey recorded from the reference and,
when /root/reference is mounted, against the live reference objects."""
import pytest
import torch

from oracle import ref_loader

PARAMS = {"mnist3": 3209153, "mnist12": 11607809, "cifar15": 14467684, "celeba20": 19207460}


def _build(name):
    import lvae_b200
    from lvae_b200.configs import baseline_kwargs
    return lvae_b200.LadderVAE(**baseline_kwargs(name)), baseline_kwargs(name)


@pytest.mark.parametrize("name", sorted(PARAMS))
def test_parameter_count_and_geometry(name):
    m, kw = _build(name)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == PARAMS[name]
    assert m.n_layers == len(kw["z_dims"])
    assert int(m.overall_downscale_factor) == 2 ** (sum(kw["downsample"]) + 1)
    padded = m.get_padded_size((5, kw["color_ch"]) + tuple(kw["img_shape"]))
    assert padded == {"mnist3": [32, 32], "mnist12": [32, 32], "cifar15": [32, 32], "celeba20": [64, 64]}[name]
    assert m.get_padded_size(tuple(kw["img_shape"])) == padded                  # (H, W) form, lvae.py:338-340
    assert m.get_top_prior_param_shape() == (1, 64, 2, 2)                        # 2 x 2 top map, 2 Z channels
    assert m.get_top_prior_param_shape(7)[0] == 7
    assert tuple(m.top_down_layers[-1].top_prior_params.shape) == (1, 64, 2, 2)
    assert m.top_down_layers[-1].top_prior_params.requires_grad                  # --learn-top-prior
    if ref_loader.reference_available():
        ref = ref_loader.load_reference()["lvae"].LadderVAE(**kw)
        assert sum(p.numel() for p in ref.parameters() if p.requires_grad) == PARAMS[name]
        assert ref.get_padded_size(tuple(kw["img_shape"])) == padded
        assert tuple(ref.get_top_prior_param_shape()) == tuple(m.get_top_prior_param_shape())
        assert int(ref.overall_downscale_factor) == int(m.overall_downscale_factor)
        assert [n for n, _ in ref.named_modules()] == [n for n, _ in m.named_modules()]     # same module tree


def test_exception_conventions_match_reference():
    m, kw = _build("mnist3")
    cases = [
        (RuntimeError, lambda mod: mod.get_padded_size((1, 2, 3))),                            # lvae.py:341-344
        (RuntimeError, lambda mod: mod.topdown_pass()),                                        # neither bu_values nor n_img_prior, :248-251
        (RuntimeError, lambda mod: mod.topdown_pass(bu_values=[None] * 3, n_img_prior=2)),     # both
        (RuntimeError, lambda mod: mod.topdown_pass(bu_values=[None] * 3, mode_layers=[0])),   # prior experiment in inference, :252-255
        (ValueError, lambda mod: mod.top_down_layers[-1](input_=torch.zeros(1, 64, 2, 2))),    # lvae_layers.py:127-128
    ]
    mods = [m]
    if ref_loader.reference_available():
        mods.append(ref_loader.load_reference()["lvae"].LadderVAE(**kw))
    for mod in mods:
        for exc, fn in cases:
            with pytest.raises(exc):
                fn(mod)
    import lvae_b200
    with pytest.raises(AssertionError):                                                        # lvae.py:60-61
        lvae_b200.LadderVAE(1, [4, 4], img_shape=(8, 8), downsample=[1], likelihood_form="bernoulli",
                            res_block_type="bacdbac", merge_type="residual")
    with pytest.raises(AssertionError):
        lvae_b200.LadderVAE(1, [4], blocks_per_layer=1, img_shape=(8, 8), downsample=[2], likelihood_form="bernoulli",
                            res_block_type="bacdbac", merge_type="residual")


def test_free_bits_matches_loader_stub():
    """The product's free_bits_kl and the stand-in the golden vectors were generated with are the same function of
    (kl, free_bits) -- including the below-threshold and batch_average branches."""
    from lvae_b200.boilr_compat import free_bits_kl, HAVE_BOILR
    if HAVE_BOILR:
        pytest.skip("real boilr installed: its own free_bits_kl is used")
    stub = ref_loader._boilr_stubs()["boilr.nn"].free_bits_kl
    g = torch.Generator().manual_seed(0)
    kl = torch.rand(6, 4, generator=g) * 3
    for fb in (0.0, 1e-7, 0.5, 2.0):
        for ba in (False, True):
            assert torch.equal(free_bits_kl(kl, fb, batch_average=ba), stub(kl, fb, batch_average=ba))
    out = free_bits_kl(kl, 1.0)
    assert tuple(out.shape) == (4,) and float(out.min()) >= 1.0
