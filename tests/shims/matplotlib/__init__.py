"""Stand-in for matplotlib (see ../README.md): utils.py:3-7 imports it at module level and the trainer's
validation pass draws spectrogram figures into TensorBoard (train_time_wi_inv.py:303-331)."""


def use(backend, **kw):
    return None
