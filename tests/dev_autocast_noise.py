"""How far do bf16-autocast gradients of the (oracle restatement of the) reference drift from fp32? (GPU, torch ops only)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from common import *
from oracle import srvp_oracle as O
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
O.USE_ATEN_LSTM = False
g = load_golden('vgg_skip_nc3')
cfg = g['cfg']
T, B = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (g['T'], g['B'])
m = build_model(cfg, g['res_gain'], 1)
sd0 = {k: v.cuda() for k, v in m.state_dict().items()}
x = make_input(T, B, cfg['nc'], 123).cuda()
torch.manual_seed(7)
rnd = O.draw_randoms(cfg, T, T, B, training=True)
rnd = {k: ([e.cuda() for e in v] if isinstance(v, list) else v.cuda()) for k, v in rnd.items()}
res = {}
for mode in ('fp32', 'bf16'):
    sdo = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16')):
        o = O.forward(sdo, cfg, x, T, g['dt'], rnd, training=True)
    l = O.elbo({k: (v.float() if torch.is_tensor(v) else v) for k, v in o.items()}, x, g['loss_cfg'])
    l[0].backward()
    res[mode] = (float(l[0]), {k: v.grad for k, v in sdo.items() if v.requires_grad})
print('loss fp32', res['fp32'][0], 'bf16', res['bf16'][0], 'rel', abs(res['fp32'][0] - res['bf16'][0]) / res['fp32'][0])
errs = sorted((rel_l2(res['bf16'][1][k], res['fp32'][1][k]), k) for k in res['fp32'][1])
print('autocast-vs-fp32 grad rel_l2: median', errs[len(errs) // 2][0], 'max', errs[-1])
for e, k in errs[::8]:
    print(f'   {k:40s} {e:.3e}')
