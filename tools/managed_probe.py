import ctypes as C, torch
torch.zeros(1).cuda()
rt = C.CDLL("libcudart.so.12")
p = C.c_void_p()
rc = rt.cudaMallocManaged(C.byref(p), C.c_size_t(2048), C.c_uint(1))
rt.cudaGetErrorString.restype = C.c_char_p
print("cudaMallocManaged rc", rc, rt.cudaGetErrorString(rc))
attr = C.c_int()
for name, a in (("managedMemory", 83), ("concurrentManagedAccess", 89), ("pageableMemoryAccess", 88)):
    rt.cudaDeviceGetAttribute(C.byref(attr), a, 0); print(name, attr.value)
