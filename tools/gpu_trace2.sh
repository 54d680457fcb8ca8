TRACE_MODE=2 bash tools/gpu_trace.sh 2>&1 | grep -A 6 "fused trace" | tail -40
