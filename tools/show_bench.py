import json,sys
d=json.load(open(sys.argv[1]))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "floor", d["e2e"]["copy_floor_ms"])
print("roofline", d["roofline"]["frac"], d["roofline"]["step_frac"], d["roofline"]["kernel_ms"])
print("parity", d["parity_check"])
print("ev", {k:v for k,v in d["ev"].items() if k!="path"})
print("cpu", (d["cpu_baseline"] or {}).get("value"), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
