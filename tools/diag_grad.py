import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import pdes_oracle as orc
from models.codec import DenseED
from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
from utils.image_gradient import SobelFilter
g = np.load('tests/golden/densenet_full32.npz')
cfg = dict(in_channels=1, out_channels=3, imsize=32, blocks=[6,8,6], growth_rate=16, init_features=48)
plan = orc.densenet_plan(**cfg)
names = [str(s) for s in g['param_names']]
res = {}
for impl in (1, 0, 3, 4, 5):
    sd = orc.make_state(plan, int(g['seed']))
    model = DenseED(1, 3, 32, [6,8,6]); model.load_state_dict(sd); model = model.cuda(); model.conv_impl = impl
    K = orc.make_input(int(g['B']), 32, int(g['seed'])).cuda()
    sob = SobelFilter(32, device='cuda')
    model.train(); model.zero_grad()
    out = model(K); out.retain_grad()
    loss = conv_constitutive_constraint(K, out, sob) + conv_continuity_constraint(out, sob)
    d, n = conv_boundary_condition(out); loss = loss + (d + n) * 10
    loss.backward(); torch.cuda.synchronize()
    params = dict(model.named_parameters())
    pos = 0; rows = []
    for i, nme in enumerate(names):
        gr = params[nme].grad.double().cpu().numpy().ravel()
        k = int(g['grads64_head_len'][i]); ref = g['grads64_head'][pos:pos+k]; pos += k
        e_head = np.linalg.norm(gr[:k]-ref)/max(np.linalg.norm(ref),1e-30)
        e_norm = abs(np.linalg.norm(gr)-g['grad_norm64'][i])/g['grad_norm64'][i]
        rows.append((e_head, e_norm, g['grad_err32'][i]/g['grad_norm64'][i]))
    rows = np.array(rows)
    o = out.detach().cpu().double().numpy()
    print('impl', impl, 'out rel', np.linalg.norm(o-g['out64'])/np.linalg.norm(g['out64']),
          'dout rel', np.linalg.norm(out.grad.cpu().double().numpy()-g['dout64'])/np.linalg.norm(g['dout64']),
          'head err median %.2e max %.2e | norm err median %.2e max %.2e | ref fp32 floor median %.2e max %.2e' % (
          np.median(rows[:,0]), rows[:,0].max(), np.median(rows[:,1]), rows[:,1].max(), np.median(rows[:,2]), rows[:,2].max()))
    worst = np.argsort(-rows[:,1])[:6]
    print('   worst norm errs:', [(names[j], '%.2e' % rows[j,1], '%.2e' % rows[j,2]) for j in worst])
