# the 1 -> 8 scaling run on ONE 8 x B200 box, launched the way the driver launches bench.py
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topology_8gpu.txt 2>&1; lscpu | head -25 >> gpurun_out/topology_8gpu.txt; numactl -H >> gpurun_out/topology_8gpu.txt 2>&1
python bench.py --gpus 1 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --no-cpu > gpurun_out/scale_n$n.json 2>> gpurun_out/scale.err
done
for n in 1 2 4 8; do python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
print($n, round(d["value"] / 1e3, 1), round(d["ms_per_step"], 3), [round(v, 2) for v in d.get("per_rank_ms", [])],
      "strong", round(d["strong_scaling"]["value"] / 1e3, 1), round(d["strong_scaling"]["ms_per_step"], 3),
      "e2e", round(d["e2e"]["value"] / 1e3, 1), "u8 e2e", round(d["device_format_u8"]["e2e"]["value"] / 1e3, 1))
PY
done
