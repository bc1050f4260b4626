"""utils/deepspeed.py of the reference.  DeepSpeed ZeRO-1 + fp16 is replaced by the native path (fp16 tensor-core
operands, fp32 master weights, one NCCL gradient all-reduce), so the config is informational and inputs stay fp32."""


def get_deepspeed_config(args):
    return {"train_batch_size": getattr(args, "effective_batch_size", None),
            "gradient_clipping": getattr(args, "max_grad_norm", 0.0),
            "fp16": {"enabled": True, "native": "lavender_b200"}, "zero_optimization": {"stage": 0}}


def fp32_to_fp16(batch):
    return batch
