import torch, time
x = torch.empty(256*28*28, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def t(fn, n=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        a.record(st); fn(); b.record(st); b.synchronize(); tot += a.elapsed_time(b)
    return tot / n * 1e3
print("single copy 802816 B: %.1f us" % t(lambda: d.copy_(x, non_blocking=True)))
streams = [torch.cuda.Stream() for _ in range(4)]
def chunks(k):
    n = x.numel() // k
    ev = torch.cuda.Event(); ev.record(st)
    evs = []
    for i in range(k):
        s = streams[i]; s.wait_event(ev)
        with torch.cuda.stream(s):
            d[i*n:(i+1)*n].copy_(x[i*n:(i+1)*n], non_blocking=True)
        e = torch.cuda.Event(); e.record(s); evs.append(e)
    for e in evs: st.wait_event(e)
print("2 chunks on 2 streams: %.1f us" % t(lambda: chunks(2)))
print("4 chunks on 4 streams: %.1f us" % t(lambda: chunks(4)))
y = torch.empty(2560, dtype=torch.float32).pin_memory(); dy = torch.empty_like(y, device="cuda")
print("10 KB copy: %.1f us" % t(lambda: dy.copy_(y, non_blocking=True)))
h = torch.zeros(2, dtype=torch.float64).pin_memory(); dd = torch.zeros(2, dtype=torch.float64, device="cuda")
t0 = time.perf_counter()
for _ in range(200): h.copy_(dd, non_blocking=True); st.synchronize()
print("D2H 16 B + sync (host wall): %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
