#!/bin/bash
# quick A/B of the SRIF / hybrid production kernels at the bench configuration (kernel ms by CUDA events)
python tools/sweep_nl_chunks.py srif 2>&1 | head -2
SWEEP_EPOCHS=200 python tools/sweep_nl_chunks.py srif hybrid 2>&1 | grep default
