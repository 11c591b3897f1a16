"""GPU bring-up probe for the tcgen05 GEMM (run under gpurun):  python tools/gpu_probe.py

A child process walks the case list and prints one RESULT line per case; if it hangs or crashes (wrong descriptor ->
illegal instruction / deadlock) the parent records that case and restarts the child after it, so one bad variant
cannot take the whole call down.  Writes gpurun_out/probe.json."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases():
    out = []
    for (m, n, k) in ((128, 256, 64), (300, 520, 200)):
        for a_mn in (False, True):
            for b_mn in (False, True):
                for bn in (128, 256, 192, 160):
                    if bn == 160 and b_mn:
                        continue
                    if m == 128 and bn not in (128, 256):
                        continue
                    out.append(dict(a_mn=a_mn, b_mn=b_mn, bn=bn, m=m, n=n, k=k, probe=None))
    # alternative UMMA descriptor byte offsets (k_lbo, k_sbo, mn_lbo, mn_sbo); evaluated only to diagnose failures
    for probe in ((0, 0, 1024, 8192), (0, 0, 8192, 8192), (0, 0, 128, 1024), (0, 0, 1024, 128), (0, 0, 16384, 1024)):
        for a_mn, b_mn in ((True, False), (False, True)):
            out.append(dict(a_mn=a_mn, b_mn=b_mn, bn=128, m=300, n=520, k=200, probe=probe, alt="mn"))
    for probe in ((0, 1024, 0, 0), (16, 512, 0, 0), (1024, 1024, 0, 0), (8192, 1024, 0, 0)):
        out.append(dict(a_mn=False, b_mn=False, bn=128, m=300, n=520, k=200, probe=probe, alt="k"))
    return out


def child(start: int):
    import torch

    import algonauts2025_b200  # noqa: F401
    from algonauts2025_b200 import ops

    cs = cases()
    for idx in range(start, len(cs)):
        c = cs[idx]
        print(f"BEGIN {idx}", flush=True)
        torch.manual_seed(0)
        m, n, k = c["m"], c["n"], c["k"]
        A = torch.randn(m, k, device="cuda").bfloat16()
        B = torch.randn(n, k, device="cuda").bfloat16()
        ref = A.float() @ B.float().t()
        def mn_op(x):
            mn, kk = x.shape
            pad = (mn + 7) // 8 * 8
            store = torch.zeros(kk, pad, device="cuda", dtype=x.dtype)
            store[:, :mn] = x.t()
            return ops.Operand(store, inner=mn, rows=kk, row_stride=pad, mn_major=True)

        a_op = mn_op(A) if c["a_mn"] else ops.kmajor(A)
        b_op = mn_op(B) if c["b_mn"] else ops.kmajor(B)
        out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float32)
        ops.gemm(a_op, b_op, out, m, n, k, ldd=n, block_n=c["bn"], probe=c["probe"])
        torch.cuda.synchronize()
        err = (out - ref).abs()
        bad = ~(err <= 1e-2 + 1e-2 * ref.abs())
        res = dict(max_err=float(err.nan_to_num(1e9).max()), frac_bad=float(bad.float().mean()), nan=int(out.isnan().sum()))
        if res["frac_bad"] > 0:
            eb = bad.float()
            res["bad_rows_128"] = [round(float(eb[i:i + 128].mean()), 3) for i in range(0, m, 128)]
            res["bad_cols_64"] = [round(float(eb[:, j:j + 64].mean()), 3) for j in range(0, n, 64)]
        print(f"RESULT {idx} " + json.dumps(res), flush=True)


def main():
    cs = cases()
    results = {}
    start, restarts = 0, 0
    while start < len(cs) and restarts < 12:
        proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", str(start)], stdout=subprocess.PIPE,
                                stderr=subprocess.PIPE, text=True)
        try:
            stdout, stderr = proc.communicate(timeout=240)
            timed_out = False
        except subprocess.TimeoutExpired:
            proc.kill()
            stdout, stderr = proc.communicate()
            timed_out = True
        last_begin = None
        for line in stdout.splitlines():
            if line.startswith("BEGIN "):
                last_begin = int(line.split()[1])
            elif line.startswith("RESULT "):
                _, idx, payload = line.split(" ", 2)
                r = json.loads(payload)
                r["status"] = "ok" if r["frac_bad"] == 0 and r["nan"] == 0 else "mismatch"
                results[int(idx)] = r
        if last_begin is not None and last_begin not in results:
            results[last_begin] = {"status": "timeout" if timed_out else "crash", "stderr": stderr[-400:]}
            start = last_begin + 1
            restarts += 1
        elif last_begin is None:
            results[start] = {"status": "child-failed", "stderr": stderr[-800:]}
            break
        else:
            start = last_begin + 1
    default_ok = all(results.get(i, {}).get("status") == "ok" for i, c in enumerate(cs) if c["probe"] is None)
    named = {}
    for i, c in enumerate(cs):
        if i not in results:
            continue
        if c["probe"] is not None and default_ok:
            continue  # alternatives are only interesting when a default failed
        key = f"{i:02d}_m{c['m']}n{c['n']}k{c['k']}_A{'mn' if c['a_mn'] else 'k'}_B{'mn' if c['b_mn'] else 'k'}_bn{c['bn']}_probe{c['probe']}"
        named[key] = results[i]
        print(key, json.dumps(results[i]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(named, f, indent=1)
    print(f"PROBE SUMMARY: defaults {'ALL OK' if default_ok else 'FAILED'}; {sum(v['status'] == 'ok' for v in named.values())}/{len(named)} listed ok")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        main()
