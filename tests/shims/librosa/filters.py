from oracle import np_oracle


def mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
    return np_oracle.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
