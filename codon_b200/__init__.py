"""codon_b200: B200 (sm_100a) engine for the CODON guided depth super-resolution forward pass.

Public surface (mirrors the reference's modules):
  codon_b200.CODON_x4 / CODON_x8 / CODON_x16 : CODONNet
  codon_b200.CAC_module                      : CAC_channel, CAC_spatial, BasicConv, ChannelPool, Flatten, logsumexp_2d
  codon_b200.attention.ResCBAM               : ChannelGate, SpatialGate, ResCBAM, ResCBAM_c, ResCBAM_d
  codon_b200.ssim_2                          : ssim_exact
  codon_b200.Loger                           : Logger
  codon_b200.test                            : main(), test(model), EvaluationResults(depth_high, output)
  codon_b200.engine                          : Engine (ctypes binding of include/codon_b200.h)
"""
__version__ = "0.1.0"
