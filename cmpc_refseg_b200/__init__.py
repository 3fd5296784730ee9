"""cmpc_refseg_b200 -- B200-native (sm_100a) implementation of the CMPC head of zigonk/CMPC-Refseg.

Public surface:
  CMPC_model.LSTM_model     drop-in for the reference's head path (same ctor kwargs / attribute names)
  head.CMPCHeadB200         host orchestration of the kernels over the C ABI (include/cmpc_b200.h)
  ops                       torch.ops.cmpc.* custom ops over the same C ABI (import cmpc_refseg_b200.ops to register)
  runner.HostPipeline / TrainPipeline   host-buffer entry points (H2D / compute overlap)
  build.build()             compiles libcmpc_b200.so with nvcc for sm_100a
"""
from . import build  # noqa: F401

__all__ = ["build"]
