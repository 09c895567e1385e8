for seg in 2048 3072 3584 4096 6144 8192; do
  python tools/gpu_diag.py 67108864 $seg 10 | tail -1 | cut -c1-200
done
