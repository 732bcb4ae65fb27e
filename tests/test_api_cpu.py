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


def test_low_precision_gradient_side_table():
    """ops._lowp_grads hands the bf16 copy a producing kernel wrote to the consumer of the SAME fp32 gradient tensor and to
    nobody else (host logic only: CPU tensors stand in for device buffers)."""
    from lvae_b200 import ops
    ops._lowp_grads.clear()
    g32 = torch.randn(2, 4, 4, 64)
    g16 = g32.bfloat16()
    ops._lowp_grad_put(g32, g16)
    view = g32.permute(0, 3, 1, 2).permute(0, 2, 3, 1)                 # what autograd hands on: a view of the same storage
    assert ops._lowp_grad_take(view, torch.bfloat16) is g16
    assert ops._lowp_grad_take(view, torch.bfloat16) is None           # collected once
    # a gradient that autograd summed with another one is a new tensor: no entry
    ops._lowp_grad_put(g32, g16)
    assert ops._lowp_grad_take(g32 + 1.0, torch.bfloat16) is None
    # the buffer was written after the copy was taken (version counter moved): the copy is stale
    g32.add_(1.0)
    assert ops._lowp_grad_take(g32, torch.bfloat16) is None
    # wrong dtype / shape on the consumer side
    ops._lowp_grad_put(g32, g16)
    assert ops._lowp_grad_take(g32, torch.float16) is None
    ops._lowp_grad_put(g32, g16[:1])
    assert ops._lowp_grad_take(g32, torch.bfloat16) is None
    # the entry keeps the fp32 tensor alive, so its address cannot be recycled while the entry exists
    ops._lowp_grad_put(g32, g16)
    ptr = g32.data_ptr()
    del g32, view
    assert ops._lowp_grads[ptr][0].data_ptr() == ptr and len(ops._lowp_grads[ptr]) == 3
    # uncollected entries are dropped wholesale, not accumulated forever
    keep = [torch.zeros(1) for _ in range(300)]
    for t in keep:
        ops._lowp_grad_put(t, t.bfloat16())
    assert len(ops._lowp_grads) <= 257
    ops._lowp_grads.clear()
