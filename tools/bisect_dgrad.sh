for shape in "32 32 128 32 32 7" "32 16 64 64 64 3" "32 8 32 128 128 3" "64 32 128 128 128 3"; do
  python tools/run_layer.py dgrad $shape | tail -1
  python tools/run_layer.py wgrad $shape | tail -1
done
