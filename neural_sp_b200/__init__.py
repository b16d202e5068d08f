"""neural_sp_b200: B200-native (sm_100a) speech-encoder + CTC / RNN-T hot path of hirofumi0810/neural_sp.

Host side mirrors the reference's ``neural_sp.models.seq2seq.{encoders,decoders}`` module API; all
arithmetic runs in hand-written CUDA behind the C ABI of ``libnsp_b200.so`` (include/nsp_b200.h).
"""
__version__ = "0.1.0"
