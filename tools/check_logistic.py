"""Dev helper: accuracy of the chain's fast logistic against mpmath-free numpy long double. (GPU)"""
import ctypes, numpy as np, subprocess, os, sys
src = r'''
#include "atlasqtl_b200/csrc/aq_common.cuh"
extern "C" __global__ void k(const double* x, double* y, int n) { int i = blockIdx.x*blockDim.x+threadIdx.x; if (i<n) y[i] = aq::logistic_neg(x[i]); }
extern "C" void run(const double* hx, double* hy, int n) { double *x,*y; cudaMalloc(&x,8*n); cudaMalloc(&y,8*n); cudaMemcpy(x,hx,8*n,cudaMemcpyHostToDevice); k<<<(n+255)/256,256>>>(x,y,n); cudaMemcpy(hy,y,8*n,cudaMemcpyDeviceToHost); cudaFree(x); cudaFree(y); }
'''
open("/tmp/lg.cu","w").write(src)
subprocess.check_call(["nvcc","-gencode","arch=compute_100a,code=sm_100a","-O3","-I.","-shared","-Xcompiler","-fPIC","-o","/tmp/lg.so","/tmp/lg.cu"])
lib = ctypes.CDLL("/tmp/lg.so")
x = np.concatenate([np.linspace(-50, 50, 200001), np.linspace(-720, 720, 20001), np.random.default_rng(0).normal(0, 5, 200000)])
y = np.empty_like(x)
lib.run(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(x)))
xl = x.astype(np.longdouble)
ref = (1 / (1 + np.exp(xl)))
rel = np.abs((y - ref) / np.maximum(ref, np.longdouble(1e-300))).astype(float)
print("max abs err", float(np.abs(y - ref).max()), "max rel err (|x|<=700)", rel[np.abs(x) <= 690].max())
