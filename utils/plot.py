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
