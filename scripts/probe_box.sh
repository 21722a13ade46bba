set -x
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA" ; nvidia-smi --query-gpu=name,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv
ls /usr/lib/x86_64-linux-gnu | grep -i nccl; python -c "import torch,os; print(torch.cuda.nccl.version()); import nvidia.nccl, glob; print(glob.glob(os.path.dirname(nvidia.nccl.__file__)+'/lib/*'))"
ulimit -l
