"""Diagnostic (GPU): UMMA throughput per operand layout / N on ONE SM and on ALL SMs at once, alone and with concurrent TMEM
loads / shared-memory stores (the fused conv kernel's situation); TMEM load throughput.  Evidence for DESIGN.md section 4."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcad_b200
lib = bcad_b200._lib.load()
out = torch.zeros(6, dtype=torch.int64, device="cuda")
SMS = torch.cuda.get_device_properties(0).multi_processor_count
def run(N, al, bl, albo, asbo, blbo, bsbo, reps=2000, ldw=4, x16=0, a_off=0, alt=0, grid=1, conc=0, stw=0, hbm=0):
    p = np.array([N, al, bl, albo, asbo, blbo, bsbo, reps, ldw, x16, a_off, alt, grid, conc, stw, hbm], np.int32)
    bcad_b200._lib.check(lib.bcad_selftest_umma_bench(C.c_void_p(p.ctypes.data), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    return o[0] / reps, o[1] / reps, o[3] / reps, o[4], o[5]
for N in (64, 128, 256):
    a = run(N, 0, 0, 2176, 128, N * 16, 128)
    b = run(N, 2, 2, 16, 1024, 16, 1024)
    c = run(N, 0, 2, 2176, 128, 16, 1024)
    d = run(N, 4, 4, 16, 512, 16, 512)
    print(f"N={N}: cycles/MMA  A,B no-swizzle {a[0]:.1f} | A,B SW128 {b[0]:.1f} | A none, B SW128 {c[0]:.1f} | A,B SW64 {d[0]:.1f}")
for a_off in (0, 16, 32, 64):
    for alt in (0, 1):
        r = run(64, 0, 0, 2176, 128, 1024, 128, 2000, 4, 0, a_off, alt)
        print(f"N=64 no-swizzle, A start +{a_off} B, alternate operands={alt}: {r[0]:.1f} cycles/MMA")
for ldw in (1,):
    for x16 in (0, 1):
        r = run(64, 2, 2, 16, 1024, 16, 1024, 2000, ldw, x16)
        print(f"tmem ld: {ldw} warps, x{16 if x16 else 32}: {r[1]:.1f} cycles per load per warp")
# one SM vs the whole chip, and the MMA stream beside the traffic the fused kernel's other warps generate
for N in (64, 128, 256):
    for grid in (1, 2, SMS // 2, SMS, 2 * SMS):
        r = run(N, 0, 0, 2176, 128, N * 16, 128, reps=20000, grid=grid, alt=1)
        print(f"N={N} grid={grid}: {r[0]:.1f} cycles/MMA on CTA 0, {r[2]:.1f} on the slowest CTA")
for grid in (1, SMS):
    for ldw, stw in ((0, 0), (4, 0), (0, 3), (4, 3)):
        r = run(64, 0, 0, 2176, 128, 1024, 128, reps=20000, ldw=ldw, grid=grid, alt=1, conc=1, stw=stw)
        print(f"N=64 grid={grid} concurrent: {ldw} TMEM-load warps, {stw * 32} storing threads: {r[0]:.1f} cycles/MMA (slowest CTA {r[2]:.1f}); "
              f"{r[3]} loads/warp, {r[4]} store rounds in that time")

# the same MMA stream beside HBM store traffic (the conv epilogue writes 2.8 TB/s while the MMAs run)
for grid in (1, SMS):
    r = run(64, 0, 0, 2176, 128, 1024, 128, reps=20000, ldw=4, grid=grid, alt=1, conc=1, stw=3, hbm=1)
    print(f"N=64 grid={grid} concurrent: 4 TMEM-load warps, 96 threads storing to HBM: {r[0]:.1f} cycles/MMA (slowest CTA {r[2]:.1f}); "
          f"{r[4]} store rounds = {r[4] * 96 * 16 * grid / (r[0] * 20000) :.0f} B/clk chip-wide")
# sustained: does the per-MMA time (in SM cycles) stay at 48 once the chip runs into its power cap?
import subprocess, time
def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception as e:
        return str(e)
for N, hbm in ((64, 0), (64, 1), (256, 0)):
    t0, i = time.time(), 0
    while time.time() - t0 < 4.0:
        r = run(N, 0, 0, 2176, 128, N * 16, 128, reps=200000, ldw=4, grid=SMS, alt=1, conc=1, stw=3, hbm=hbm)
        if i % 40 == 0:
            print(f"sustained N={N} hbm_stores={hbm} t={time.time() - t0:.2f}s: {r[0]:.1f} cycles/MMA (slowest CTA {r[2]:.1f}) | sm MHz, W, power cap: {smi()}")
        i += 1
