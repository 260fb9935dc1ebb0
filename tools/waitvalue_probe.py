# diagnose stream wait-value support (driver API through cuda-python) with runtime-created objects
from cuda import cuda, cudart
print("init", cudart.cudaSetDevice(0), cudart.cudaFree(0))
err, s = cudart.cudaStreamCreateWithFlags(cudart.cudaStreamNonBlocking)
err, p = cudart.cudaMalloc(128)
print("memset", cudart.cudaMemset(p, 0, 128))
print("wait GEQ 0 on runtime stream/ptr:", cuda.cuStreamWaitValue32(s, p, 0, 0))
print("wait GEQ 0 at +4:", cuda.cuStreamWaitValue32(s, p + 4, 0, 0))
print("sync", cudart.cudaStreamSynchronize(s))
for ver in (11070, 12000, 12090):
    r = cudart.cudaGetDriverEntryPointByVersion(b"cuStreamWaitValue32", ver, cudart.cudaEnableDefault) if hasattr(cudart, "cudaGetDriverEntryPointByVersion") else None
    print("entry point by version", ver, r)
print("entry point default", cudart.cudaGetDriverEntryPoint(b"cuStreamWaitValue32", cudart.cudaEnableDefault))
