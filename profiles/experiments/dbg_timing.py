import os, sys, ctypes as C
os.environ["DUO_B200_LIB"] = os.path.join(os.getcwd(), "scratch/lib_dbg.so")
sys.argv = ["bench_tc.py"]
exec(open("scratch/bench_tc.py").read())
import numpy as np
buf = np.zeros(8192, dtype=np.int64)
lib.duo_debug_read.argtypes = [C.c_void_p, C.c_size_t]
rc = lib.duo_debug_read(buf.ctypes.data, buf.nbytes)
s0 = buf[0:256].reshape(64, 4); s1 = buf[1024:1280].reshape(64, 4); m = buf[2048:2304].reshape(64, 4)
t0 = s0[8, 0]
print("rc", rc)
print("slot0: [wait_start, S_ready, P_stored, arrived] rel; per-tile period")
for j in range(8, 20):
    print(j, (s0[j] - t0).tolist(), "| slot1", (s1[j] - t0).tolist(), "| mma [pre-wait P0, got P0, pre-wait P1, got P1]", (m[j] - t0).tolist())
per = np.diff(s0[8:60, 1]); print("slot0 period mean", per.mean(), "softmax busy (S_ready->arrived) mean", (s0[8:60,3]-s0[8:60,1]).mean(), "wait for S mean", (s0[8:60,1]-s0[8:60,0]).mean())
per = np.diff(s1[8:60, 1]); print("slot1 period mean", per.mean(), "softmax busy mean", (s1[8:60,3]-s1[8:60,1]).mean(), "wait for S mean", (s1[8:60,1]-s1[8:60,0]).mean())
print("phase offset slot1-slot0 S_ready mean", (s1[8:60,1]-s0[8:60,1]).mean())
print("mma: wait for P0 mean", (m[8:60,1]-m[8:60,0]).mean(), "wait for P1 mean", (m[8:60,3]-m[8:60,2]).mean(), "issue PV0+S0' mean", (m[8:60,2]-m[8:60,1]).mean())
f = buf[4096:4096+512].reshape(64, 8)
print("mma fine: [after v_full, after issue_pv, after commit, after k_full, after issue_s, after commit] relative to got-P0")
for j in range(8, 14):
    print(j, (f[j, :6] - m[j, 1]).tolist())
d = f[8:60]
print("means: v_full wait", (d[:,0]-m[8:60,1]).mean(), "issue_pv", (d[:,1]-d[:,0]).mean(), "commit", (d[:,2]-d[:,1]).mean(), "k_full wait", (d[:,3]-d[:,2]).mean(), "issue_s", (d[:,4]-d[:,3]).mean(), "commit", (d[:,5]-d[:,4]).mean())
