"""Per-shape GEMM timing inside one eager training step pair (mlm + sap): where does magic_gemm time go?"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import magic_b200
from magic_b200 import _lib, ops
from magic_b200.graph_index import batch_to_device
from magic_b200.train_step import PretrainStepper

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "magic_s_pretrain_b64"]
dev = torch.device("cuda", 0)
cfg_s, cfg_t = bench.make_cfgs(w, 0.1)
torch.manual_seed(1)
student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(torch.bfloat16)
stepper = PretrainStepper(student, None)
pools = {t: [batch_to_device(b, dev) for b in bench.make_pool(t, 2, w, 1234)] for t in ("mlm", "sap")}
for i in range(4):
    stepper.step("mlm" if i % 2 == 0 else "sap", pools["mlm" if i % 2 == 0 else "sap"][0])
torch.cuda.synchronize()
_lib.profile_start()
_lib.call("magic_delay", int(60e6), _lib.stream())
for i in range(4):
    stepper.step("mlm" if i % 2 == 0 else "sap", pools["mlm" if i % 2 == 0 else "sap"][1])
torch.cuda.synchronize()
prof = _lib.profile_stop()
agg = collections.defaultdict(lambda: [0.0, 0])
for e0, e1, a in prof["magic_gemm"]:
    key = (a[11], a[12], a[13], "A" + ("k" if a[3] == 1 else "m"), "B" + ("k" if a[6] == 1 else "n"),
           "a%d b%d c%d" % (a[1], a[5], a[9]), "lda%d ldb%d" % (max(a[2], a[3]), max(a[6], a[7])))
    agg[key][0] += e0.elapsed_time(e1)
    agg[key][1] += 1
tot = sum(v[0] for v in agg.values())
print("total gemm ms over 4 steps: %.2f" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    M, N, K = k[0], k[1], k[2]
    print("%7.3f ms  n=%3d  avg %7.1f us  %6.1f TF  %s" % (v[0], v[1], 1e3 * v[0] / v[1], 2.0 * M * N * K * v[1] / (v[0] * 1e-3) / 1e12, k))
for name, recs in sorted(prof.items(), key=lambda kv: -sum(e0.elapsed_time(e1) for e0, e1, _ in kv[1])):
    print("%-28s %8.3f ms  %d calls" % (name, sum(e0.elapsed_time(e1) for e0, e1, _ in recs), len(recs)))
