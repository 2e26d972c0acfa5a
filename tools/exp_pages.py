"""Experiment: does the outermost pass of 512^3 speed up on memory mapped through the virtual-memory API with its
recommended granularity (larger GPU pages -> fewer TLB entries for the 512 rows a tile touches)?"""
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

try:
    from cuda.bindings import driver as cu
except ImportError:
    from cuda import cuda as cu

st = torch.cuda.current_stream().cuda_stream
s.init(0)
n = 512 ** 3
P = 512 * 512
nbytes = 16 * n
res = {}


def ck(r):
    if r[0] != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(r[0]))
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)


def ev(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)


prop = cu.CUmemAllocationProp()
prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
prop.location.id = 0
gmin = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
grec = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED))
res["granularity_min"], res["granularity_recommended"] = int(gmin), int(grec)


def vmm_alloc(size, align):
    size = (size + align - 1) // align * align
    h = ck(cu.cuMemCreate(size, prop, 0))
    ptr = ck(cu.cuMemAddressReserve(size, align, 0, 0))
    ck(cu.cuMemMap(ptr, size, 0, h, 0))
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(cu.cuMemSetAccess(ptr, size, [acc], 1))
    return int(ptr)


y = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
res["torch_tensor_x_pass"] = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 1, 512, P, False, 1.0, st))
res["torch_tensor_y_pass"] = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st))
for name, align in (("vmm_recommended", int(grec)), ("vmm_512MB_aligned", 512 << 20), ("vmm_1GB_single_chunk_aligned", 1 << 30)):
    try:
        p = vmm_alloc(nbytes, align)
        torch.cuda.synchronize()
        ck(cu.cuMemcpyDtoD(p, y.data_ptr(), nbytes))
        res[name + "_x_pass"] = ev(lambda: s.fft_axis(p, p, 1, 512, P, False, 1.0, st))
        res[name + "_y_pass"] = ev(lambda: s.fft_axis(p, p, 512, 512, 512, False, 1.0, st))
    except Exception as e:  # noqa: BLE001
        res[name] = "failed: " + str(e)[:100]
print(json.dumps(res))
