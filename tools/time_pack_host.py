"""Throughput of the host-side GSO mask packer (magat_gso_pack_host) on this box's cores."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magat_pathplanning_b200", "lib", "libmagat_gat.so"))
L.magat_gso_pack_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
B, N = 256, 1000
S = torch.zeros(B, 1, N, N).pin_memory()
S[:, :, ::7, ::11] = 0.25
W = (N + 31) // 32
bits = torch.zeros(B * N, W, dtype=torch.int32).pin_memory()
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for nt in (1, 2, 4, 8, 16, 32, 64):
    L.magat_gso_pack_host(S.data_ptr(), 0, B * N, N, bits.data_ptr(), nt)
    t0 = time.perf_counter()
    for _ in range(3):
        L.magat_gso_pack_host(S.data_ptr(), 0, B * N, N, bits.data_ptr(), nt)
    dt = (time.perf_counter() - t0) / 3
    print(f"threads {nt:3d}: {S.numel() * 4 / dt / 1e9:7.1f} GB/s  ({dt * 1e3:.1f} ms per {S.numel() * 4 / 1e9:.2f} GB)")
if torch.cuda.is_available():
    d = torch.empty_like(S, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        d.copy_(S, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"H2D pinned: {S.numel() * 4 / dt / 1e9:.1f} GB/s")
