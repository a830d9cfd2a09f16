"""B200-native implementation of the DCASE2019-task4 baseline hot path.

waveform -> log-mel -> CRNN forward/backward -> mean-teacher step, behind the reference's Python surface
(``models.CRNN.CRNN``, ``DataLoad`` transforms, ``utils.Scaler``, ``main.train``), with all arithmetic in
hand-written sm_100a CUDA kernels reached through the C ABI of ``include/dcase_b200.h``.
"""
__version__ = "0.1.0"
