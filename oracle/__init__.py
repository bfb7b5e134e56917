"""CPU oracle for the time-domain vocoding hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and there only as the checker or
the CPU baseline -- never from the product package
(``neural-vocoders-as-speech-enhancers_b200``).

Two restatements of the reference algorithm live here:

* ``np_oracle``   -- float64 numpy, written from the maths, the
                     parity checker.
* ``torch_port``  -- fp32 functional-PyTorch port that issues the same library calls
                     the reference issues on CPU (``torch.stft``, ``F.conv1d`` ...);
                     it is the CPU baseline that ``bench.py`` times (kind = "port").

Pinning: the reference ships no golden vectors.  Both restatements
are pinned against outputs of the reference's own code run in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference`` by file path and writes
``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks them.  The one piece
that cannot be pinned that way is ``librosa.filters.mel`` (librosa==0.10.2.post1 is a
pip dependency of the reference and is not installed here): its published algorithm
is restated in ``np_oracle.mel_filterbank`` and cross-checked against
``torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")`` --
"parity unpinned" for that function only.
"""
