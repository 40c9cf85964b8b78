timeout 600 python -m pytest tests -q -m gpu -x -k "conv_backward or train_step_chain or bn_silu or adamw or detection_loss" 2>&1 | tail -25
