for kb in 11 8 11 8; do
  echo -n "kb $kb: "
  MSIG_EPI_MIN_KB=$kb timeout 120 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['inference']['ms_per_batch'])"
done
