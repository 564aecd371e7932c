mkdir -p gpurun_out
for v in 0 1 0 1; do echo "== RN_GRAPH_PDL=$v"; RN_GRAPH_PDL=$v RN_GRAPH_DEBUG=1 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps|graph|recnow" | cut -c1-220; done | tee gpurun_out/pdl2.txt
echo "== plain launches (RN_GRAPH=0)"; RN_GRAPH=0 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call" | cut -c1-100
RN_GRAPH_PDL=1 python -m pytest tests/test_pairwise_gpu.py tests/test_launch_and_peer_gpu.py tests/test_listwise_pairs_gpu.py -m gpu -x -q 2>&1 | tail -2
