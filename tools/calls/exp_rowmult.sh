for m in 1 2 3 8; do
  GDN_EW_ROW_MULT=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mult',$m,'img/s',round(d['value'],1),'ms',round(d['ms_per_step'],2), d['clocks'])"
done
