for f in 0 65536 131072 262144 524288 1048576 589824; do
  echo "flags=$f"; SKY_DEBUG_FLAGS=$f python tools/run_layer.py da 32 32 128 32 32 7 2>&1 | tail -1
done
echo plain32to3; python tools/run_layer.py plain 32 32 128 32 3 7 1 | tail -1
echo plain64to32k3; python tools/run_layer.py plain 32 32 128 64 32 3 1 | tail -1
echo da32to32k3; python tools/run_layer.py da 32 32 128 32 32 3 | tail -1
echo dense; python tools/run_layer.py dense 32 4096 4096 | tail -1; python tools/run_layer.py dense 32 8192 4096 | tail -1;  python tools/run_layer.py dense 32 4096 8192 | tail -1
