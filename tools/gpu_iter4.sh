#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python tools/fa_ab.py --streams 16 --reps 3 2>&1 | tee gpurun_out/fa_ab.txt | tail -4)
(timeout 400 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum -k regex:flash_attn --csv --log-file gpurun_out/fa_ab_launches.csv python tools/fa_ab.py --streams 16 --reps 1 > gpurun_out/fa_ab_ncu.log 2>&1; echo "ncu list exit $?"; python - <<'PY'
import csv, io, collections
lines=[l for l in open("gpurun_out/fa_ab_launches.csv") if not l.startswith("==")]
agg=collections.defaultdict(list)
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r.get("Metric Name")=="gpu__time_duration.sum":
        v=float(r["Metric Value"].replace(",","")); u=r.get("Metric Unit","ns")
        agg[r["Kernel Name"][:40]].append(v*{"ns":1e-3,"us":1,"ms":1e3}.get(u,1e-3))
for k,v in agg.items(): print(k, len(v), "launches, avg %.1f us, min %.1f, max %.1f" % (sum(v)/len(v), min(v), max(v)))
PY
)
(timeout 400 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:flash_attn_kernel -s 3 -c 1 -o gpurun_out/prof_flash_r2b -f python tools/fa_ab.py --streams 16 --reps 1 > gpurun_out/prof_flash_r2b.log 2>&1; echo "ncu full exit $?"; python tools/summarize_ncu.py report gpurun_out/prof_flash_r2b.ncu-rep gpurun_out/prof_flash_r2b.md 2>&1 | tail -30)
