f() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], round(d['gcups']), 'ms',round(d['ms'],2),'fwd',round(d['fwd_ms'],2),'tb',round(d['tb_ms'],2),'launches',d['launches'],'bad',d['mismatching'])
    elif 'Error' in l or 'error' in l: print(l.strip())
"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg4 or phases or cfg3 or cfg2" 2>&1 | tail -4
echo "== default"; python tools/run_configs.py --scale 1 --check 2000 --steps 5 2>&1 | f
echo "== default scale 0.4"; python tools/run_configs.py --scale 0.4 --check 500 --steps 5 --only cfg4 2>&1 | f
echo "== phases 1 scale 0.4"; UBU_TB_PHASES=1 python tools/run_configs.py --scale 0.4 --check 500 --steps 5 --only cfg4 2>&1 | f
echo "== phases 1"; UBU_TB_PHASES=1 python tools/run_configs.py --scale 1 --check 300 --steps 5 --only cfg4 2>&1 | f
