python tools/state_probe.py 3 cfg4; python tools/state_probe.py 3 cfg2
for lib in "" variants/libubu_sw_036208b.so; do
  UBU_SW_LIB=${lib:+$PWD/$lib} python bench.py --no-configs --no-multi --no-cpu 2>/dev/null | tail -1 > gpurun_out/e2e_ab.json
  python - "$lib" <<PY
import json,sys; d=json.loads(open("gpurun_out/e2e_ab.json").read()); print(sys.argv[1] or "current", round(d["value"]), "e2e", round(d["e2e"]["value"]), "one_call", round(d["e2e"]["one_call"]["value"]), "from_text", round(d["e2e"]["from_text"]["value"]))
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
