for v in s2m4c4 s2m2c2 s1m4c4 s3m3c3 s2m3c3; do
  if [ -z "$v" ]; then unset BDLRU_LIB; else export BDLRU_LIB=$PWD/datamining_recblr_b200/variants/libbdlru_conv1d_$v.so; fi
  python tools/conv_bench.py 8192 200 256 bf16
  python tools/conv_bench.py 2048 50 128 bf16
done
unset BDLRU_LIB
python tools/conv_bench.py 8192 200 256 f32
python -m pytest tests -m gpu -q -p no:cacheprovider -k "conv" 2>&1 | tail -2
