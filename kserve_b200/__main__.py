"""`python -m kserve_b200 --model_dir /mnt/models --model_name llama` — the entrypoint that replaces
`python -m huggingfaceserver --backend huggingface` (python/huggingfaceserver/huggingfaceserver/__main__.py:60-345)
for Llama-family decoders: same --model_id/--model_dir/--model_name/--max_model_len/--dtype flags, backend fixed to
the B200 engine (no vLLM / transformers dispatch), plus --tensor_parallel_size (one process per GPU under torchrun)."""
import argparse
import os
import sys

from .kserve_api import model_server
from .kserve_api.model_server import ModelServer


def main(argv=None):
    parser = argparse.ArgumentParser(parents=[model_server.parser])
    parser.add_argument("--model_dir", default="/mnt/models", help="HF checkpoint directory (config.json, *.safetensors, tokenizer)")
    parser.add_argument("--model_id", default=None, help="alias of --model_dir (no Hub access in this runtime)")
    parser.add_argument("--max_model_len", "--max_length", type=int, default=None)
    parser.add_argument("--dtype", default="auto", choices=["auto", "bfloat16"], help="the B200 engine computes in bf16")
    parser.add_argument("--backend", default="b200", choices=["b200"])
    parser.add_argument("--max_batch", type=int, default=32)
    parser.add_argument("--continuous_batching", action="store_true",
                        help="iteration-level batching: requests join / leave the running decode batch between steps")
    parser.add_argument("--tensor_parallel_size", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    parser.add_argument("--enable_batcher", action="store_true", help="batch V1 :predict like the Go agent (--max-batchsize/--max-latency)")
    parser.add_argument("--max-batchsize", dest="max_batchsize", type=int, default=32)
    parser.add_argument("--max-latency", dest="max_latency", type=int, default=5000)
    args, _ = parser.parse_known_args(argv)
    from .generative_model import B200GenerativeModel
    path = args.model_id or args.model_dir
    nccl_id, rank = None, int(os.environ.get("RANK", "0"))
    if args.tensor_parallel_size > 1:
        import torch
        import torch.distributed as dist
        from .tp import broadcast_nccl_id
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("gloo")
        nccl_id = broadcast_nccl_id(rank)
    model = B200GenerativeModel(args.model_name, path, max_model_len=args.max_model_len, max_batch=args.max_batch,
                                device=int(os.environ.get("LOCAL_RANK", "0")), tensor_parallel_size=args.tensor_parallel_size,
                                tp_rank=rank, nccl_id=nccl_id, continuous_batching=args.continuous_batching)
    model.load()
    if rank != 0:
        from .tp import follower_loop
        return follower_loop(model)
    ModelServer(http_port=args.http_port, enable_latency_logging=args.enable_latency_logging, grpc_port=args.grpc_port,
                enable_grpc=args.enable_grpc).start([model])


if __name__ == "__main__":
    sys.exit(main())
