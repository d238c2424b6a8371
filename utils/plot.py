"""Stats / figure writers with the reference's names (utils/plot.py:17-94, 261-273 upstream).
Figures are written only when a real matplotlib is importable; the numeric .txt/.npy outputs
are always written.  `save_stats` tolerates 1-epoch runs (a 0-d loadtxt result crashes the
upstream version)."""
import numpy as np

from .misc import to_numpy


def _pyplot():
    try:
        import matplotlib
        if getattr(matplotlib, "__pdes_shim__", False):
            return None
        import matplotlib.pyplot as plt
        plt.switch_backend('agg')
        return plt
    except Exception:
        return None


def plot_prediction_det(save_dir, target, prediction, epoch, index, plot_fn='contourf', cmap='jet',
                        same_scale=False, row_labels=None, col_labels=None):
    target, prediction = to_numpy(target), to_numpy(prediction)
    np.save(save_dir + '/pred_epoch{}_{}.npy'.format(epoch, index), np.stack([target, prediction]))
    plt = _pyplot()
    if plt is None:
        return
    rows = [target, prediction, target - prediction]
    fig, axes = plt.subplots(3, target.shape[0], figsize=(3.5 * target.shape[0], 9))
    axes = np.atleast_2d(axes)
    for r, fields in enumerate(rows):
        for c in range(target.shape[0]):
            ax = axes[r, c]
            if plot_fn == 'contourf':
                im = ax.contourf(fields[c], 50, cmap=cmap)
            else:
                im = ax.imshow(fields[c], cmap=cmap, origin='lower', interpolation='bilinear')
            ax.set_axis_off()
            fig.colorbar(im, ax=ax, fraction=0.046, pad=0.04)
    fig.savefig(save_dir + '/pred_epoch{}_{}.png'.format(epoch, index), bbox_inches='tight')
    plt.close(fig)


def plot_prediction_det_animate2(save_dir, target, prediction, epoch, index, i_plot, plot_fn='imshow',
                                 cmap='jet', same_scale=False):
    """Frame writer of the solver's --animate option (utils/plot.py upstream; imported unconditionally by
    solve_conv_mixed_residual.py:23): the numeric frame is always saved, the figure when matplotlib exists."""
    target, prediction = to_numpy(target), to_numpy(prediction)
    np.save(save_dir + '/frame{:04d}_epoch{}_{}.npy'.format(int(i_plot), epoch, index), np.stack([target, prediction]))
    plot_prediction_det(save_dir, target, prediction, epoch, index, plot_fn=plot_fn, cmap=cmap, same_scale=same_scale)


def plot_prediction_bayes2(save_dir, target, pred_mean, pred_var, epoch, index, plot_fn='imshow', cmap='jet',
                           same_scale=False):
    """Predictive mean / variance writer of train_cglow_reverse_kl.py:207-209 (utils/plot.py:181 upstream): the
    numeric arrays are always saved, a 4-row figure (target, mean, error, variance) when matplotlib exists."""
    target, pred_mean, pred_var = to_numpy(target), to_numpy(pred_mean), to_numpy(pred_var)
    np.save(save_dir + '/pred_bayes_epoch{}_{}.npy'.format(epoch, index), np.stack([target, pred_mean, pred_var]))
    plt = _pyplot()
    if plt is None:
        return
    rows = [target, pred_mean, target - pred_mean, pred_var]
    fig, axes = plt.subplots(4, target.shape[0], figsize=(3.5 * target.shape[0], 12))
    axes = np.atleast_2d(axes)
    for r, fields in enumerate(rows):
        for c in range(target.shape[0]):
            ax = axes[r, c]
            im = (ax.contourf(fields[c], 50, cmap=cmap) if plot_fn == 'contourf' else
                  ax.imshow(fields[c], cmap=cmap, origin='lower', interpolation='bilinear'))
            ax.set_axis_off()
            fig.colorbar(im, ax=ax, fraction=0.046, pad=0.04)
    fig.savefig(save_dir + '/pred_bayes_epoch{}_{}.png'.format(epoch, index), bbox_inches='tight')
    plt.close(fig)


def save_samples(save_dir, images, epoch, index, name, nrow=4, heatmap=True, cmap='jet', title=False):
    """Sample-grid writer of train_cglow_reverse_kl.py:215-216 (utils/plot.py:644 upstream): (N, C, H, W) samples,
    one grid per channel; the array is always saved."""
    images = to_numpy(images)
    np.save(save_dir + '/{}_epoch{}_{}.npy'.format(name, epoch, index), images)
    plt = _pyplot()
    if plt is None:
        return
    n, ch = images.shape[0], images.shape[1]
    ncol = (n + nrow - 1) // nrow
    for c in range(ch):
        fig, axes = plt.subplots(nrow, ncol, figsize=(2.5 * ncol, 2.5 * nrow))
        for k, ax in enumerate(np.atleast_1d(axes).ravel()):
            ax.set_axis_off()
            if k < n:
                ax.imshow(images[k, c], cmap=cmap if heatmap else 'gray', origin='lower', interpolation='bilinear')
        fig.savefig(save_dir + '/{}_epoch{}_{}_c{}.png'.format(name, epoch, index, c), bbox_inches='tight')
        plt.close(fig)


def save_stats(save_dir, logger, *metrics):
    plt = _pyplot()
    for metric in metrics:
        arr = np.atleast_1d(np.asarray(logger[metric], dtype=np.float64))
        np.savetxt(save_dir + f'/{metric}.txt', arr)
        if plt is None or arr.size == 0:
            continue
        arr2 = arr.reshape(arr.shape[0], -1)
        lines = plt.plot(range(1, arr2.shape[0] + 1), arr2)
        plt.legend(lines, [f'{arr2[-5:, i].mean():.4f}' for i in range(arr2.shape[1])])
        plt.savefig(save_dir + f'/{metric}.pdf')
        plt.close()
