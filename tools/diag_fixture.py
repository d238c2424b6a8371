"""Per-tensor gradient error of a DenseED fixture for several conv_impl values (GPU):
    python tools/diag_fixture.py densenet_bottleneck32 [impl ...]
Prints, per implementation, the noise-floor ratios the parity test uses and where (in named_parameters order)
the tensors leave the floor - a ReLU-mask flip shows as everything UPSTREAM of one layer moving together."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pdes_oracle as orc
from models.codec import DenseED
from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
from utils.image_gradient import SobelFilter
name = sys.argv[1]
impls = [int(a) for a in sys.argv[2:]] or [1, 0, 6, 7]
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', name + '.npz'))
cfg = dict(in_channels=int(g["cfg_in_channels"]), out_channels=int(g["cfg_out_channels"]), imsize=int(g["cfg_imsize"]),
           blocks=[int(b) for b in g["cfg_blocks"]], growth_rate=int(g["cfg_growth_rate"]), init_features=int(g["cfg_init_features"]))
bn = int(g["bn_size"]) if "bn_size" in g.files else 0
plan = orc.densenet_plan(**cfg, bottleneck=bn)
names = [str(s) for s in g['param_names']]
for impl in impls:
    sd = orc.make_state(plan, int(g['seed']))
    model = DenseED(cfg["in_channels"], cfg["out_channels"], cfg["imsize"], cfg["blocks"], growth_rate=cfg["growth_rate"],
                    init_features=cfg["init_features"], **(dict(bottleneck=True, bn_size=bn) if bn else {}))
    model.load_state_dict(sd); model = model.cuda(); model.conv_impl = impl
    K = orc.make_input(int(g['B']), cfg["imsize"], int(g['seed'])).cuda()
    sob = SobelFilter(cfg["imsize"], device='cuda')
    if os.environ.get("DIAG_EVAL_FIRST"):   # the parity test's order: an eval-mode forward before the training step
        model.eval()
        with torch.no_grad():
            model(K)
    model.train(); model.zero_grad()
    out = model(K); out.retain_grad()
    loss = conv_constitutive_constraint(K, out, sob) + conv_continuity_constraint(out, sob)
    d, n = conv_boundary_condition(out); loss = loss + (d + n) * 10
    loss.backward(); torch.cuda.synchronize()
    params = dict(model.named_parameters())
    pos = 0; ratios = []
    for i, nme in enumerate(names):
        gr = params[nme].grad.double().cpu().numpy().ravel()
        if "grads64" in g.files:
            ref = g["grads64"][pos:pos + gr.size]; pos += gr.size
            err = np.linalg.norm(gr - ref)
        else:
            k = int(g['grads64_head_len'][i]); ref = g['grads64_head'][pos:pos+k]; pos += k
            err = np.linalg.norm(gr[:k]-ref) * np.sqrt(gr.size / k)
        nrm = float(g['grad_norm64'][i])
        ratios.append(err / max(3 * float(g['grad_err32'][i]), 1e-5 * nrm))
    ratios = np.array(ratios)
    o = out.detach().cpu().double().numpy()
    print('impl', impl, 'out rel %.2e' % (np.linalg.norm(o-g['out64'])/np.linalg.norm(g['out64'])),
          'median ratio %.2f  frac<=1 %.2f  max %.1f' % (np.median(ratios), np.mean(ratios <= 1), ratios.max()))
    last_bad = max([i for i, r in enumerate(ratios) if r > 1.0], default=-1)
    print('   last tensor above the bar: %d of %d (%s); ratios by quarter of the network: %s' % (
        last_bad, len(names), names[last_bad] if last_bad >= 0 else '-',
        ['%.2f' % np.median(q) for q in np.array_split(ratios, 8)]))
