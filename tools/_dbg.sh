for d in 0 1 2 4 8 16 3 11 15 31; do EVREP_TAF_DBG=$d python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_dbg$d.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/b_dbg$d.json'));print($d, d['ms_per_step'],d['roofline']['kernel_ms'])"; done
