"""One-cycle learning-rate schedule (host scalar math; reference utils/practices.py:16-41)."""
import math


class OneCycleScheduler(object):
    """Linear warm-up from lr_max/div_factor to lr_max over the first `pct_start` of training,
    then cosine annealing down to lr_max/div_factor/1e4."""

    def __init__(self, lr_max, div_factor=25., pct_start=0.3):
        self.lr_max = lr_max
        self.div_factor = div_factor
        self.pct_start = pct_start
        self.lr_low = lr_max / div_factor

    def step(self, pct):
        if pct <= self.pct_start:
            return self.lr_low + (pct / self.pct_start) * (self.lr_max - self.lr_low)
        t = (pct - self.pct_start) / (1 - self.pct_start)
        end = self.lr_low / 1e4
        return end + (self.lr_max - end) / 2 * (math.cos(math.pi * t) + 1)


def adjust_learning_rate(optimizer, lr):
    for group in optimizer.param_groups:
        group['lr'] = lr
    return lr


def find_lr(net, trn_loader, optimizer, loss_fn, weight_bound, init_value=1e-8, final_value=10., beta=0.98,
            device='cuda:0'):
    """Exponential LR range test (only referenced from commented-out code upstream)."""
    num = max(len(trn_loader) - 1, 1)
    mult = (final_value / init_value) ** (1 / num)
    lr, avg, best, log_lrs, losses = init_value, 0., 0., [], []
    adjust_learning_rate(optimizer, lr)
    for i, (inp,) in enumerate(trn_loader, start=1):
        inp = inp.to(device)
        optimizer.zero_grad()
        loss = loss_fn(inp, net(inp), weight_bound)
        avg = beta * avg + (1 - beta) * loss.item()
        smooth = avg / (1 - beta ** i)
        if i > 1 and smooth > 4 * best:
            break
        if smooth < best or i == 1:
            best = smooth
        losses.append(smooth)
        log_lrs.append(math.log10(lr))
        loss.backward()
        optimizer.step()
        lr *= mult
        adjust_learning_rate(optimizer, lr)
    return log_lrs, losses
