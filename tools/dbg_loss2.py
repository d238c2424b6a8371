import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from models.codec import DenseED
from pde_surrogate_b200.engine import TrainStep
torch.manual_seed(1)
dev = torch.device("cuda:0")
model = DenseED(1, 3, 64, [6, 8, 6]).to(dev)
ts = TrainStep(model)
g = torch.Generator(device="cpu").manual_seed(1)
Ks = [torch.exp(0.5 * torch.randn(32, 1, 64, 64, generator=g)).to(dev) for _ in range(4)]
def st(tag):
    torch.cuda.synchronize()
    f = ts.flat
    print(tag, "flat abs-sum %.4f nan %d | gflat abs-sum %.4e | m %.3e v %.3e | l4 %s | hyper %s" % (
        float(f.abs().sum()), int(torch.isnan(f).sum()), float(ts.gflat.abs().sum()), float(ts.m.abs().sum()),
        float(ts.v.abs().sum()), ts.l4.tolist(), ts.hyper_dev.tolist()), flush=True)
st("init")
ts.capture(Ks[0])
st("after capture")
for i in range(3):
    loss = ts.step_graph(Ks[i % 4])
    st("replay %d loss %.4f staticK %.3f" % (i, float(loss), float(ts.static_K.sum())))
model.eval()
with torch.no_grad():
    o = model(Ks[0])
print("eval out abs-sum", float(o.abs().sum()))
