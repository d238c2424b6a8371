"""Minimal stand-in for matplotlib.pyplot, used ONLY when the real package is missing: the reference
scripts call a handful of pyplot functions at module level / at the end of a run
(train_codec_mixed_residual.py:33, solve_conv_mixed_residual.py:31-32, 180-184).  Every call is a no-op."""


class _Noop(object):
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


_noop = _Noop()


def switch_backend(*args, **kwargs):
    return None


def __getattr__(name):   # plt.imshow / plt.colorbar / plt.savefig / plt.close / plt.figure ...
    if name.startswith("__"):
        raise AttributeError(name)
    return _noop
