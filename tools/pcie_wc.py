"""Does write-combined pinned memory change the copy floor of the host path?  (cudaHostAllocWriteCombined for the input,
the output, or both; one GPU, the byte counts of one bench step, both directions at once.)"""
import ctypes, json, time
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaStreamCreate.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
n_out, n_in = 2147483648, 910869458
def halloc(n, flags):
    p = ctypes.c_void_p(); assert rt.cudaHostAlloc(ctypes.byref(p), n, flags) == 0; ctypes.memset(p, 1, n); return p
def dalloc(n):
    p = ctypes.c_void_p(); assert rt.cudaMalloc(ctypes.byref(p), n) == 0; return p
s1, s2 = ctypes.c_void_p(), ctypes.c_void_p(); rt.cudaStreamCreate(ctypes.byref(s1)); rt.cudaStreamCreate(ctypes.byref(s2))
d_out, d_in = dalloc(n_out), dalloc(n_in)
res = {}
for name, fin, fout in [("pinned/pinned", 0, 0), ("wc input", 4, 0), ("wc output", 0, 4), ("wc both", 4, 4)]:
    h_in, h_out = halloc(n_in, fin), halloc(n_out, fout)
    best = 1e9
    for it in range(4):
        rt.cudaDeviceSynchronize(); t = time.perf_counter()
        rt.cudaMemcpyAsync(h_out, d_out, n_out, 2, s1)
        rt.cudaMemcpyAsync(d_in, h_in, n_in, 1, s2)
        rt.cudaDeviceSynchronize(); dt = time.perf_counter() - t
        if it: best = min(best, dt)
    res[name] = round(best * 1e3, 2)
    rt.cudaFreeHost(h_in); rt.cudaFreeHost(h_out)
print(json.dumps({"ms_both_directions": res, "bytes": {"h2d": n_in, "d2h": n_out}}))
