"""Throughput of ONE beam-search call as a function of the number of captions stacked along the row axis
(captions are independent units: a stacked decode is bit-identical to separate decodes, tests/test_gpu_parity.py::
test_properties_at_full_size), optionally with several calls in flight on their own streams.

    python tools/stack_probe.py [b1,b2,...] [lanes]      e.g.  100,200,300,400,600,800 1

Prints one JSON line per (b, lanes): ms per decode, captions/s, per-phase ms (profiled eager run)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def inputs(b, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    D, L, R = 50, 10, 20
    det = torch.relu(torch.randn((b, D, 2048), device=dev, generator=g))
    ds = torch.relu(torch.randn((b, L, R, 2048), device=dev, generator=g))
    nv = torch.randint(1, R + 1, (b, L), device=dev, generator=g)
    ds = ds * (torch.arange(R, device=dev)[None, None, :] < nv[:, :, None]).unsqueeze(-1)
    verbs = -torch.ones((b, L), dtype=torch.float64, device=dev)
    verbs[:, 2] = 17
    return det, ds, verbs


def main():
    from models import ControllableCaptioningModel
    bs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "100,200,300,400,600,800").split(",")]
    lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    iters = 6
    dev = "cuda:0"
    torch.manual_seed(1234)
    models = [ControllableCaptioningModel(20, 10000, 2, verb_tables=({}, {})).to(dev).eval()]
    for _ in range(lanes - 1):
        m2 = ControllableCaptioningModel(20, 10000, 2, verb_tables=({}, {})).to(dev).eval()
        m2.load_state_dict(models[0].state_dict())
        models.append(m2)
    streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
    for b in bs:
        stat = [inputs(b, 7 + i, dev) for i in range(lanes)]
        for _ in range(3):
            for m, s in zip(models, stat):
                m.beam_search_v(s, [3, -1], 5, 1, gt=True)
        torch.cuda.synchronize()
        cur = torch.cuda.current_stream()
        rounds = []
        for _ in range(5):                      # five timed rounds: the median is robust against clock / power transients
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            for st in streams:
                st.wait_event(e0)
            for i in range(iters * lanes):
                with torch.cuda.stream(streams[i % lanes]):
                    models[i % lanes].beam_search_v(stat[i % lanes], [3, -1], 5, 1, gt=True)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                cur.wait_event(ev)
            e1.record(cur)
            torch.cuda.synchronize()
            rounds.append(e0.elapsed_time(e1) / (iters * lanes))
        ms = sorted(rounds)[len(rounds) // 2]
        out = {"b": b, "lanes": lanes, "ms_per_decode": ms, "ms_min": min(rounds), "ms_max": max(rounds), "captions_per_s": b / ms * 1e3}
        if lanes == 1:
            eng = models[0]._eng
            eng.set_profiling(True)
            models[0].beam_search_v(stat[0], [3, -1], 5, 1, gt=True)
            out["phases_ms"] = {n: round(t, 4) for n, t, _ in eng.phase_times() if t > 0}
            eng.set_profiling(False)
        print(json.dumps(out), flush=True)
        del stat
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
