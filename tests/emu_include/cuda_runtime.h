// stub: tests/cuda_emu.h stands in for the CUDA toolkit when device headers are compiled for the host emulator
