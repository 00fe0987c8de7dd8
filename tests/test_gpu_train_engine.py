"""GPU parity of the whole training step on libhulk_sm100 kernels (TrainEngine, SURVEY.md §8 f1 / BASELINE config 4)
against the torch fp32 autograd graph over the same parameters -- i.e. what the reference's train.py:33-35 executes
(train-mode BatchNorm, K-row head == 1000-row head on the live channels, sigmoid, .double(), nn.BCELoss, backward)."""
import warnings

import pytest
import torch

import hulk_keypoints_b200 as hk
from hulk_keypoints_b200 import train_ops

warnings.filterwarnings("ignore")
pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def synth(gen, B, H, W, K=4):
    uv = torch.stack([torch.randint(8, W - 8, (B, K), generator=gen), torch.randint(8, H - 8, (B, K), generator=gen)], -1).float()
    img = torch.rand(B, 3, H, W, generator=gen)
    return img.cuda(), uv.cuda()


def cosine(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.fixture(scope="module")
def pair():
    """Engine step and autograd step on identical models and one batch."""
    torch.manual_seed(3)
    m = hk.KeypointsGauss(4).cuda().train()
    ref = hk.KeypointsGauss(4).cuda().train()
    ref.load_state_dict(m.state_dict())
    img, uv = synth(torch.Generator().manual_seed(11), 2, 64, 96)
    eng = m.train_engine(2, 64, 96)
    loss = eng.forward_backward(img, uv=uv).clone()
    grads = {n: eng.grad(p).clone() for n, p in m.named_parameters()}
    loss_ref = train_ops.sigmoid_bce_loss(ref.forward_logits(img), uv=uv, sigma=8.0)
    loss_ref.backward()
    return m, ref, eng, img, uv, loss, grads, loss_ref


def discs(gen, B, H, W, K=4):
    """Images with a coloured disc at each keypoint + noise, and the (x, y) labels (SURVEY.md §8c F-trn)."""
    uv = torch.stack([torch.randint(8, W - 8, (B, K), generator=gen), torch.randint(8, H - 8, (B, K), generator=gen)], -1).float()
    img = 0.2 * torch.rand(B, 3, H, W, generator=gen)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    colors = torch.tensor([[1.0, 0.1, 0.1], [0.1, 1.0, 0.1], [0.1, 0.1, 1.0], [1.0, 1.0, 0.1]])
    for b in range(B):
        for k in range(K):
            disc = ((xx - uv[b, k, 0]) ** 2 + (yy - uv[b, k, 1]) ** 2 <= 36).float()
            img[b] = img[b] * (1 - disc) + 0.8 * disc * colors[k].view(3, 1, 1) + 0.2 * img[b] * disc
    return img.cuda(), uv.cuda()


def test_engine_loss_matches_autograd_on_raw_init(pair):
    """Untrained weights + noise images: the loss agrees; the gradient there is a numerically hostile quantity (rounding only the
    INPUT of the fp32 reference to bf16 already moves it to cosine ~0.83), so directions are gated on the trained
    fixture below and only magnitudes and the structural zeros are checked here."""
    m, ref, eng, img, uv, loss, grads, loss_ref = pair
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    for name, p in ref.named_parameters():
        g, r = grads[name], p.grad
        if r.norm() == 0:
            assert g.norm() == 0, name
        else:
            assert 0.7 < float(g.norm() / r.norm()) < 1.4, name
    # the 996 dead rows of the 1000-row scoring conv receive exactly zero gradient, as in the reference
    assert grads["resnet.resnet34_8s.fc.weight"][4:].abs().sum().item() == 0.0
    assert grads["resnet.resnet34_8s.fc.bias"][4:].abs().sum().item() == 0.0


def test_engine_training_converges_and_gradients_match_autograd_on_trained_weights():
    """120 steps of our own train step (TrainEngine + FusedAdam, zero-copy gradient hand-off) on synthetic discs, then the
    gradient of a fresh batch against torch fp32 autograd on the same weights: loss within 1e-3, every one of the 110
    parameter gradients within cosine 0.98 (measured: >= 0.993, mean 0.9990)."""
    from hulk_keypoints_b200.optim import FusedAdam
    torch.manual_seed(0)
    B, H, W = 4, 128, 160
    m = hk.KeypointsGauss(4).cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    gen = torch.Generator().manual_seed(42)
    losses = []
    for _ in range(120):
        img, uv = discs(gen, B, H, W)
        losses.append(train_ops.train_step(m, opt, img, uv, sigma=8.0).item())
    assert losses[-1] < 0.1 * losses[0], (losses[0], losses[-1])
    eng = m.train_engine(B, H, W)
    assert opt.flat_grad.data_ptr() == eng.flat_grad.data_ptr()       # the engine writes where the optimiser reads
    ref = hk.KeypointsGauss(4).cuda().train()
    ref.load_state_dict(m.state_dict())
    img, uv = discs(gen, B, H, W)
    loss = eng.forward_backward(img, uv=uv).item()
    loss_ref = train_ops.sigmoid_bce_loss(ref.forward_logits(img), uv=uv, sigma=8.0)
    loss_ref.backward()
    assert abs(loss - loss_ref.item()) < 1e-3 * abs(loss_ref.item()), (loss, loss_ref.item())
    rows = []
    for (name, p), (_, q) in zip(ref.named_parameters(), m.named_parameters()):
        if p.grad.norm() == 0:
            assert eng.grad(q).norm() == 0, name
            continue
        rows.append((cosine(eng.grad(q), p.grad), float(eng.grad(q).norm() / p.grad.norm()), name))
    rows.sort()
    for c, ratio, name in rows[:5]:
        print(f"cos {c:.5f} |g|/|ref| {ratio:.4f} {name}")
    assert rows[0][0] > 0.98, rows[:5]
    assert sum(c for c, _, _ in rows) / len(rows) > 0.997
    assert all(0.95 < ratio < 1.05 for _, ratio, _ in rows), [r for r in rows if not 0.95 < r[1] < 1.05][:5]


def test_engine_updates_batchnorm_buffers_like_torch(pair):
    m, ref = pair[0], pair[1]
    sm, sr = m.state_dict(), ref.state_dict()
    for k in sr:
        if k.endswith("num_batches_tracked"):
            assert int(sm[k]) == int(sr[k]) == 1, k
        elif "running_" in k:
            assert torch.allclose(sm[k], sr[k], rtol=2e-2, atol=2e-3), (k, (sm[k] - sr[k]).abs().max().item())


def test_engine_is_deterministic_and_graph_replay_equals_eager(pair):
    m, _, eng, img, uv, loss, grads, _ = pair
    l2 = eng.forward_backward(img, uv=uv).clone()          # CUDA-graph replay
    assert l2.item() == loss.item()
    assert torch.equal(eng.flat_grad, torch.cat([torch.nn.functional.pad(grads[n].flatten(), (0, (-grads[n].numel()) % 4))
                                                 for n, _ in m.named_parameters()]))
    eng.use_cuda_graph = False
    l3 = eng.forward_backward(img, uv=uv).clone()          # eager launches
    eng.use_cuda_graph = True
    assert l3.item() == loss.item()


def test_dropin_loss_backward_through_engine():
    """train.py:18-26,35 unmodified: loss = forward(sample, model) on (img, fp64 gt_gauss); loss.backward() fills p.grad."""
    torch.manual_seed(5)
    m = hk.KeypointsGauss(4).cuda().train()
    img, uv = synth(torch.Generator().manual_seed(12), 2, 64, 96)
    gt = torch.stack([hk.gauss_2d_batch(96, 64, 8, uv[b, :, 0], uv[b, :, 1]) for b in range(2)])   # (B,K,H,W) fp64, dataset.py:36-44
    loss = train_ops.loss_from_batch((img.cpu(), gt), m, use_cuda=True)
    assert loss.dtype == torch.float64 and loss.requires_grad
    loss.backward()
    eng = m.train_engine(2, 64, 96)
    for p in m.parameters():
        assert p.grad is not None and torch.equal(p.grad, eng.grad(p))
    # same loss as the on-the-fly-target path (targets are bit-equal to gauss_2d_batch)
    l_uv = eng.forward_backward(img, uv=uv).item()
    assert abs(l_uv - loss.item()) < 1e-12 * abs(l_uv)


def test_unmodified_train_py_step_runs_on_the_engine():
    """The literal sequence of reference train.py:18-26,33-36 -- pred = model.forward(img).double(); nn.BCELoss()(pred, gt);
    loss.backward(); optimizer.step() -- goes through the TrainEngine (model.forward in train mode is engine-backed) and gives the
    same loss and gradients as the fused path."""
    torch.manual_seed(6)
    m = hk.KeypointsGauss(4).cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-4)
    img, uv = synth(torch.Generator().manual_seed(13), 2, 64, 96)
    gt = torch.stack([hk.gauss_2d_batch(96, 64, 8, uv[b, :, 0], uv[b, :, 1]) for b in range(2)])
    opt.zero_grad()
    pred = m.forward(img).double()
    assert pred.requires_grad and pred.shape == (2, 4, 64, 96)
    loss = torch.nn.BCELoss()(pred, gt)
    loss.backward()
    eng = m.train_engine(2, 64, 96)
    split = {n: p.grad.clone() for n, p in m.named_parameters()}
    fused_loss = eng.forward_backward(img, uv=uv).item()
    assert abs(fused_loss - loss.item()) < 1e-6 * abs(fused_loss)
    for n, p in m.named_parameters():
        g = eng.grad(p)
        if g.norm() == 0:
            assert split[n].norm() == 0, n
        else:
            assert cosine(split[n], g) > 0.9999, (n, cosine(split[n], g))
    before = m.resnet.resnet34_8s.layer4[2].conv2.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, m.resnet.resnet34_8s.layer4[2].conv2.weight)
    # eval() still serves inference from the updated weights
    assert m.eval()(img).shape == (2, 4, 64, 96)


def test_two_phase_step_equals_single_phase(pair):
    """The data-parallel step runs backward in two graphs (head+layer4+layer3, then layer2+layer1+stem) so the all-reduce of the
    first part overlaps the second; gradients and loss are bit-identical to the one-graph step."""
    m, _, eng, img, uv, loss, grads, _ = pair
    ref_flat = eng.flat_grad.clone()
    eng.flat_grad.zero_()
    l = eng.forward_backward_late(img, uv).clone()
    assert eng.flat_grad[: eng.late_offset].abs().sum().item() == 0.0        # early gradients not produced yet
    assert torch.equal(eng.flat_grad[eng.late_offset:], ref_flat[eng.late_offset:])
    eng.backward_early()
    assert l.item() == loss.item() and torch.equal(eng.flat_grad, ref_flat)
    names = [n for n, _ in m.named_parameters()]
    first_late = names.index("resnet.resnet34_8s.layer3.0.conv1.weight")
    assert eng.late_offset == sum((p.numel() + 3) // 4 * 4 for p in list(m.parameters())[:first_late])


def test_engine_odd_shape_more_keypoints():
    """B=3, 72x104 (multiples of 8 only: ragged 4x16 / 8x16 tiles everywhere, TMA clipping and zero fill), K=7 (two head groups)."""
    torch.manual_seed(8)
    m = hk.KeypointsGauss(7).cuda().train()
    ref = hk.KeypointsGauss(7).cuda().train()
    ref.load_state_dict(m.state_dict())
    # a few steps so the weights are not the hostile raw init
    from hulk_keypoints_b200.optim import FusedAdam
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    gen = torch.Generator().manual_seed(21)
    for _ in range(40):
        img, uv = discs(gen, 3, 72, 104, K=4)
        uv7 = torch.cat([uv, uv[:, :3]], 1)
        train_ops.train_step(m, opt, img, uv7, sigma=8.0)
    ref.load_state_dict(m.state_dict())
    img, uv = discs(gen, 3, 72, 104, K=4)
    uv7 = torch.cat([uv, uv[:, :3]], 1)
    eng = m.train_engine(3, 72, 104)
    loss = eng.forward_backward(img, uv=uv7).item()
    loss_ref = train_ops.sigmoid_bce_loss(ref.forward_logits(img), uv=uv7, sigma=8.0)
    loss_ref.backward()
    assert abs(loss - loss_ref.item()) < 2e-3 * abs(loss_ref.item()), (loss, loss_ref.item())
    cs = [cosine(eng.grad(q), p.grad) for (n, p), (_, q) in zip(ref.named_parameters(), m.named_parameters()) if p.grad.norm() > 0]
    assert min(cs) > 0.95 and sum(cs) / len(cs) > 0.99, (min(cs), sum(cs) / len(cs))
    assert eng.grad(m.resnet.resnet34_8s.fc.weight)[7:].abs().sum().item() == 0.0
