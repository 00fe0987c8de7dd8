"""Engine vs torch-autograd gradient agreement per parameter (GPU box).  usage: diag_train_parity.py B H W"""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hulk_keypoints_b200 as hk
from hulk_keypoints_b200 import train_ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, H, W = (int(v) for v in sys.argv[1:4])
TRAIN = int(sys.argv[4]) if len(sys.argv) > 4 else 0
torch.manual_seed(3)
m = hk.KeypointsGauss(4).cuda().train()
gen = torch.Generator().manual_seed(11)
def discs(gen):
    uv = torch.stack([torch.randint(8, W - 8, (B, 4), generator=gen), torch.randint(8, H - 8, (B, 4), generator=gen)], -1).float()
    img = 0.2 * torch.rand(B, 3, H, W, generator=gen)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    colors = torch.tensor([[1.0, 0.1, 0.1], [0.1, 1.0, 0.1], [0.1, 0.1, 1.0], [1.0, 1.0, 0.1]])
    for b in range(B):
        for k in range(4):
            disc = ((xx - uv[b, k, 0]) ** 2 + (yy - uv[b, k, 1]) ** 2 <= 36).float()
            img[b] = img[b] * (1 - disc) + 0.8 * disc * colors[k].view(3, 1, 1) + 0.2 * img[b] * disc
    return img.cuda(), uv.cuda()
if TRAIN:
    from hulk_keypoints_b200.optim import FusedAdam
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    for i in range(TRAIN):
        img, uv = discs(gen)
        l = train_ops.train_step(m, opt, img, uv, sigma=8.0)
        if i % 20 == 0 or i == TRAIN - 1: print("train", i, l.item())
    img, uv = discs(gen)
else:
    uv = torch.stack([torch.randint(8, W - 8, (B, 4), generator=gen), torch.randint(8, H - 8, (B, 4), generator=gen)], -1).float().cuda()
    img = torch.rand(B, 3, H, W, generator=gen).cuda()
ref = hk.KeypointsGauss(4).cuda().train()
ref.load_state_dict(m.state_dict())
eng = m.train_engine(B, H, W)
loss = eng.forward_backward(img, uv=uv).clone()
loss_ref = train_ops.sigmoid_bce_loss(ref.forward_logits(img), uv=uv, sigma=8.0)
loss_ref.backward()
# a second fp32 reference with a tiny input perturbation shows how ill-conditioned each gradient is by itself
ref2 = hk.KeypointsGauss(4).cuda().train()
ref2.load_state_dict(m.state_dict())
l2 = train_ops.sigmoid_bce_loss(ref2.forward_logits(img.to(torch.bfloat16).float()), uv=uv, sigma=8.0)
l2.backward()
print("loss", loss.item(), "ref", loss_ref.item())
def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300))
p2 = dict(ref2.named_parameters())
for (name, p), (_, q) in zip(ref.named_parameters(), m.named_parameters()):
    g = eng.grad(q)
    if p.grad.norm() == 0: continue
    print(f"{cos(g, p.grad):8.5f} ratio {float(g.norm()/p.grad.norm()):7.4f}  ref-vs-bf16input-ref {cos(p2[name].grad, p.grad):8.5f}  |ref| {float(p.grad.norm()):.3e}  {name}")
