# e2e throughput of fp_curfit_batch_c (host pointers) for several first-chunk fractions of the two workers
for f in "0.5 0.5" "0.25 0.5" "0.25 0.35" "0.35 0.35" "0.15 0.3" "0.35 0.6"; do
  set -- $f
  FPB_HOST_FIRST_FRAC=$1 FPB_HOST_FIRST_FRAC1=$2 python bench.py --no-splev --no-shared --no-regrid --no-parcur --no-percur --no-cpu --steps 3 --warmup 3 2>/dev/null > gpurun_out/e2e_scan.json
  python - gpurun_out/e2e_scan.json $1 $2 <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("first_frac", sys.argv[2], sys.argv[3], "dev %.3f M/s" % (d["value"] / 1e6), "e2e %.3f M/s" % (d["e2e"]["value"] / 1e6), "%.1f ms" % d["e2e"]["ms_per_step"])
PY
done
