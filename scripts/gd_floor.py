"""CPU: numerical floor of the generator-alone parity test (tests/test_gpu_model.py::test_generator_forward_eval) — the
oracle against itself with (1) only the conv operands rounded to bf16 (an ideal bf16-operand / fp32-storage kernel),
(2) the product's storage rounding points, (3) round 1's rounding points.  Measured: 9.87e-3 / 1.046e-2 / 1.065e-2."""
import os
import sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import irfd_oracle as O
from parity_util import oracle_noise, _r
torch.set_num_threads(8)
torch.manual_seed(O.WEIGHT_SEED)
ref = O.IRFDRef(); O.perturb_noise_weights(ref.Gd); gd = ref.Gd.eval()
feat = torch.randn(2, 6144, generator=torch.Generator().manual_seed(O.DATA_SEED)).abs() * 0.5
oracle_noise(gd, 21)
with torch.no_grad(): img0 = gd(feat)
# ideal: bf16 operands only
hooks=[]
for blk in gd.synthesis.layers:
    for conv in (blk.conv1, blk.conv2):
        conv.weight.data = _r(conv.weight.data)
        hooks.append(conv.register_forward_pre_hook(lambda m, i: (_r(i[0]),)))
oracle_noise(gd, 21)
with torch.no_grad(): img1 = gd(feat)
print('ideal bf16-operand floor', O.rel_l2(img1,img0))
# + block outputs rounded for res>32 (product now), + to_rgb input rounding
for i,blk in enumerate(gd.synthesis.layers):
    if 2**(i+3) > 32: hooks.append(blk.register_forward_hook(lambda m,i,o:_r(o)))
oracle_noise(gd, 21)
with torch.no_grad(): img2 = gd(feat)
print('product rounding points (split<=32)', O.rel_l2(img2,img0))
for i,blk in enumerate(gd.synthesis.layers):
    if 2**(i+3) <= 32: hooks.append(blk.register_forward_hook(lambda m,i,o:_r(o)))
oracle_noise(gd, 21)
with torch.no_grad(): img3 = gd(feat)
print('round-1 rounding points (all block outputs bf16)', O.rel_l2(img3,img0))
