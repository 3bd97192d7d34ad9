"""One nm_loss_backward call of the lego configuration, for `ncu --metrics gpu__time_duration.sum` launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nerfmeshes_b200 as nm
from train_bench import CFG

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model = nm.NeRFModel(CFG).cuda().train()
g = torch.Generator().manual_seed(0)
o = torch.tensor([0.0, 0.0, 4.0]).cuda()
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1).cuda()
target = torch.rand(R, 3, generator=g).cuda()
eng = model._engine()
eng.zero_grad()
for _ in range(2):
    eng.loss_backward(o, d, 2.0, 6.0, target, training=True, seed=1)
torch.cuda.synchronize()
