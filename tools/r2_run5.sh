timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 tools/pcie_probe_ranks.py > gpurun_out/pcie_probe_n4.txt 2>&1
grep -E "^\{|Error|error" gpurun_out/pcie_probe_n4.txt | tail -8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 tools/pcie_probe_ranks.py > gpurun_out/pcie_probe_n2.txt 2>&1
grep -E "^\{" gpurun_out/pcie_probe_n2.txt | tail -3
timeout 100 python tools/pcie_probe_ranks.py > gpurun_out/pcie_probe_n1.txt 2>&1
grep -E "^\{" gpurun_out/pcie_probe_n1.txt | tail -3
