// stub: the host emulation build (warp_shim.h) supplies what dexr_kernels.cuh needs from the CUDA runtime header
