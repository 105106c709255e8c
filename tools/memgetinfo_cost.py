import time, torch, ctypes
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
f, t = ctypes.c_size_t(), ctypes.c_size_t()
x = torch.empty(20 << 30, dtype=torch.uint8, device="cuda")
for _ in range(3):
    t0 = time.perf_counter()
    for _ in range(10): rt.cudaMemGetInfo(ctypes.byref(f), ctypes.byref(t))
    print("cudaMemGetInfo ms", (time.perf_counter() - t0) * 100)
