for v in "" m54 m65 m83; do
  if [ -z "$v" ]; then unset BDLRU_LIB; else export BDLRU_LIB=$PWD/datamining_recblr_b200/variants/libbdlru_add_ln_$v.so; fi
  python tools/addln_bench.py 1638400 128 bf16 0.2
  python tools/addln_bench.py 1638400 128 bf16 0.0
  python tools/addln_bench.py 409600 64 bf16 0.2
  python tools/addln_bench.py 409600 64 f32 0.2
done
unset BDLRU_LIB
python -m pytest tests -m gpu -q -p no:cacheprovider -k "add_dropout or add_ln or layernorm" 2>&1 | tail -3
