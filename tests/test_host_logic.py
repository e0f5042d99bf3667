"""CPU suite: the host-side mirror of the reference interface (block classes, state_dict layout, convert())."""
import pytest
import torch

import srb200
from srb200 import models as M
from oracle import torch_ref as R
from oracle import ref_import

CASES = [("srcnn", (3, 64)), ("espcn", (3, 64, 4)), ("fsrcnn", (3, 4, 56, 12, 4)), ("vdsr", (3, 64, 18)),
         ("edsr", (3, 64, 16)), ("srgan_g", (3, 64, 16)), ("srgan_d", (3, 16, 32))]


@pytest.mark.parametrize("name,args", CASES)
def test_state_dict_layout_matches_reference(name, args):
    ours = M.MODELS[name](*args)
    ref = R.build(name, args, init=False)
    a, b = ours.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
    # checkpoints interchange both ways
    ours.load_state_dict(b)
    ref.load_state_dict(ours.state_dict())


@pytest.mark.parametrize("name,args", CASES)
def test_reference_init_applies_to_our_blocks(name, args):
    """utils.weights_init_* dispatch on class names; our parameter holders keep the torch classes."""
    ours = M.MODELS[name](*args)
    torch.manual_seed(0)
    {"vdsr": R.init_kaiming, "fsrcnn": R.init_fsrcnn}.get(name, R.init_normal)(ours)
    ref = R.build(name, args, seed=None if False else 0, init=False)
    torch.manual_seed(0)
    {"vdsr": R.init_kaiming, "fsrcnn": R.init_fsrcnn}.get(name, R.init_normal)(ref)
    for (k, v), (_, u) in zip(ours.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(v, u), k


def test_block_defaults_match_reference():
    b = srb200.ConvBlock(4, 8)
    assert b.conv.kernel_size == (4, 4) and b.conv.stride == (2, 2) and b.conv.padding == (1, 1)
    assert b.norm == "batch" and isinstance(b.bn, torch.nn.BatchNorm2d) and isinstance(b.act, torch.nn.ReLU)
    r = srb200.ResnetBlock(8)
    assert r.conv1.kernel_size == (3, 3) and r.conv1.padding == (1, 1)
    assert sum(1 for _ in r.modules() if isinstance(_, torch.nn.BatchNorm2d)) == 1  # one shared bn (base_networks.py:117)
    p = srb200.PSBlock(8, 4, 2)
    assert p.conv.out_channels == 16 and p.ps.upscale_factor == 2
    u = srb200.Upsample2xBlock(8, 8)
    assert isinstance(u.upsample, srb200.DeconvBlock)


def test_cpu_tensors_are_rejected_not_silently_computed():
    blk = srb200.ConvBlock(3, 4, 3, 1, 1, norm=None)
    with pytest.raises(RuntimeError, match="no CPU path"):
        blk(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        srb200.prelu(torch.zeros(4), torch.tensor([0.25]))


def test_convert_shares_parameters():
    ref = R.build("fsrcnn", (3, 4, 56, 12, 4))
    before = dict(ref.named_parameters())
    keys = list(ref.state_dict().keys())
    net = srb200.convert(ref)
    assert list(net.state_dict().keys()) == keys
    after = dict(net.named_parameters())
    assert all(after[k] is before[k] for k in before)
    assert isinstance(net.first_part, srb200.ConvBlock)
    assert isinstance(net.mid_part[5], srb200.PReLU)
    assert isinstance(net.last_part, srb200.ConvTranspose2d)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")
def test_convert_on_live_reference_and_module_shim():
    import sys
    mods = ref_import.load()
    torch.manual_seed(0)
    net = mods["edsr"].Net(3, 16, 2)
    keys = list(net.state_dict().keys())
    srb200.convert(net)
    assert list(net.state_dict().keys()) == keys
    assert isinstance(net.residual_layers[0], srb200.ResnetBlock)
    assert isinstance(net.upscale4x[0].upsample, srb200.PSBlock)
    # the star-import surface the model files rely on (srcnn.py:5,13)
    from srb200 import base_networks as B
    for n in ("torch", "ConvBlock", "PSBlock", "ResnetBlock", "DeconvBlock", "Upsample2xBlock", "DenseBlock"):
        assert n in B.__all__ and hasattr(B, n)


def test_host_optimizers_match_reference_hyperparameters():
    """srb200.host.make_optimizer restates srcnn.py:79, espcn.py:79, fsrcnn.py:106, vdsr.py:89-90, edsr.py:93; on CPU
    parameters it must not ask for the CUDA-only fused/capturable implementations."""
    from srb200 import host
    for name in ("srcnn", "espcn", "fsrcnn", "vdsr", "edsr"):
        p1 = [torch.nn.Parameter(torch.zeros(3))]
        p2 = [torch.nn.Parameter(torch.zeros(3))]
        ours = host.make_optimizer(name, p1, lr=1e-4, capturable=True)
        ref = R.make_optimizer(name, p2, lr=1e-4)
        assert type(ours) is type(ref), name
        a, b = ours.param_groups[0], ref.param_groups[0]
        for k in ("lr", "momentum", "weight_decay", "betas", "eps", "nesterov"):
            assert a.get(k) == b.get(k), (name, k)
        assert not a.get("fused") and not a.get("capturable"), name


def test_loss_selection_and_cpu_rejection():
    from srb200 import host
    import torch.nn.functional as TF
    assert host.loss_for("edsr") is TF.l1_loss and host.loss_for("espcn") is TF.mse_loss
    assert host.loss_for("edsr", fused=True) is srb200.l1_loss and host.loss_for("vdsr", fused=True) is srb200.mse_loss
    with pytest.raises(RuntimeError, match="no CPU path"):
        srb200.mse_loss(torch.zeros(4), torch.zeros(4))


def test_public_api_surface():
    for name in ("conv2d", "conv_transpose2d", "prelu", "mse_loss", "l1_loss", "convert", "GradBucket", "TrainStepGraphs",
                 "ConvBlock", "PSBlock", "ResnetBlock", "DeconvBlock", "Upsample2xBlock", "DenseBlock", "set_math",
                 "set_fuse_relu_backward", "launch_count"):
        assert hasattr(srb200, name), name


def test_prepare_marks_last_conv_block_before_a_dense_block():
    """srgan.Discriminator flattens with `.view` (srgan.py:75); prepare()/convert() make the last conv
    block hand back NCHW memory so the reference forward runs unchanged."""
    import srb200
    net = srb200.models.SRGANDiscriminator(3, 8, 32)
    flags = [b.nchw_out for b in net.conv_blocks]
    assert flags == [False] * 6 + [True] and not net.input_conv.nchw_out
    g = srb200.models.SRGANGenerator(3, 8, 1)
    assert not any(getattr(m, "nchw_out", False) for m in g.modules())  # no DenseBlock: nothing to mark


def test_convert_prepares_the_live_reference_discriminator():
    from oracle import ref_import
    if not ref_import.available():
        import pytest
        pytest.skip("reference tree not present")
    import srb200
    mods = ref_import.load()
    d = mods["srgan"].Discriminator(3, 8, 32)
    srb200.convert(d)
    assert type(d.conv_blocks[-1]).__module__ == "srb200.base_networks" and d.conv_blocks[-1].nchw_out
    assert not d.conv_blocks[0].nchw_out


def test_grad_bucket_written_flags_cpu():
    """begin_step() resets the per-step 'written' flags of direct parameters; all_reduce() zeroes direct slots nobody wrote."""
    import torch
    import srb200
    m = torch.nn.Sequential(torch.nn.Conv2d(2, 2, 3), torch.nn.PReLU())
    b = srb200.GradBucket(m, world_size=1)
    w = m[0].weight
    b.direct_ids.add(id(w))  # CPU tensors are never 'direct'; emulate a CUDA conv parameter
    w._srb_written = True
    b.begin_step()
    assert w._srb_written is False
    w.grad.fill_(3.0)
    b.all_reduce()  # nobody wrote w in this step
    assert float(w.grad.abs().sum()) == 0.0 and w._srb_written is True
    b.detach()
