#!/bin/bash
# gpurun_out/<tag>_* (tools/gpu_profile.sh) -> the text files committed under profiles/ + profiles/traffic.json
tag=${1:-r2z}; g=gpurun_out; p=profiles
cp $g/${tag}_bench.json $p/${tag}_bench_c3.json
cp $g/${tag}_bench_reference.json $p/${tag}_bench_c3_reference.json
python profiles/summarize.py launches $g/${tag}_launches.csv > $p/${tag}_c3_13band_launches.txt
units=$(python -c "import re;print(4*int(re.search(r'(\d+)',open('$g/${tag}_units.txt').read()).group(1)))")
echo "unknown-bands per captured launch: $units"
for k in update2 update2_x_left_alone direction2 rbw_down_L0 rbw_up_L0 rbw_down_L1 rbw_up_L1; do
  python profiles/summarize.py raw $g/${tag}_full_$k.ncu-rep > $p/${tag}_full_k_${k}_4band_raw.txt
done
python - "$tag" "$units" <<'PY'
import json, subprocess, sys
tag, units = sys.argv[1], sys.argv[2]
out = {"workload": "c3", "note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch captured with ncu --set full on 4 bands of the c3 "
       "workload (tools/gpu_profile.sh), divided by the unknown-bands that launch processed (level-0 kernels: 4 x the unknowns of a band)"}
for key, name in (("k_update2", "update2"), ("k_update2_x_left_alone", "update2_x_left_alone"), ("k_direction2", "direction2"),
                  ("k_rb_down", "rbw_down_L0"), ("k_rb_up", "rbw_up_L0")):
    r = subprocess.run([sys.executable, "profiles/summarize.py", "traffic", f"gpurun_out/{tag}_full_{name}.ncu-rep", units], capture_output=True, text=True)
    d = json.loads(r.stdout)
    d["capture"] = f"profiles/{tag}_full_k_{name}_4band_raw.txt"
    out[key] = d
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
for k, v in out.items():
    if isinstance(v, dict): print(k, v["kernel"], round(v["dram_bytes_per_unit"], 2), "B per unknown-band")
PY
