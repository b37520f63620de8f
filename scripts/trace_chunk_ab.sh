for c in 0 131072 262144 524288 1048576; do echo "chunk $c"; TRAY_CUDA_PIPE_CHUNK=$c python scripts/trace_e2e.py 2>&1 | grep -E "^n (2073600|8294400)"; done
