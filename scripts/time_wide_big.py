import sys; sys.argv=["x"]
exec(open("scripts/wide_time.py").read().split("only_tc =")[0])
for B in (37888, 151552):
    y0 = torch.randn(B, 64, device=dev)
    with torch.no_grad():
        tt = timeit(lambda: gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"}))
    print("B=%d %.1f us %.1f TFLOP/s" % (B, tt, B*15*262144/tt*1e-6))
