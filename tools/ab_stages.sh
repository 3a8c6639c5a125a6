for st in 2 3; do
TSG_STAGES=$st python - <<PY
import sys, torch, os
sys.path.insert(0, ".")
import __graft_entry__ as ge
t = ge.load(); t.lib(); torch.cuda.set_device(0); t.use_torch_stream()
from tools import secondary as sec
for (M,K,N,num,den) in ((4096,4096,4096,1,10),(8192,4096,14336,1,3),(4096,4096,4096,1,2)):
    Wd = t.gen_ternary(K, N, 42, num, den); W = t.DeviceTcsc.from_dense(Wd)
    Xs = [t.gen_uniform((M, K), 43 + i) for i in range(3)]; B = t.gen_uniform((N,), 44); Ys = [torch.empty((M, N), device="cuda") for _ in range(3)]
    for order in (t.ORDER_BIAS_LAST, t.ORDER_FAST):
        t.profile_enable(True); t.profile_read()
        i = [0]
        def call():
            i[0] += 1
            W.gemm(Xs[i[0] % 3], B, Ys[i[0] % 3], a=0.2, use_prelu=True, order=order)
        ms = sec._time_calls(torch, call, 12 if M == 4096 else 4)
        kms, kn = t.profile_read(); t.profile_enable(False)
        print("stages", os.environ["TSG_STAGES"], M, K, N, den, "order", order, "kernel ms %.4f" % (kms / max(kn, 1)), W.stream_info() if order == 1 else "")
    W.destroy()
PY
done
