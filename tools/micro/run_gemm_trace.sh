mkdir -p gpurun_out
T=${TAG:-r03i}
( timeout 60 tools/micro/gemm_trace 18928 3840 1152 1 0 0 10      # ViT qkv shape (bias only; the LN fold needs the producer's statistics)
  timeout 60 tools/micro/gemm_trace 18928 4304 1152 1 1 0 10      # ViT fc1 shape + GELU-tanh
  timeout 60 tools/micro/gemm_trace 18928 1152 1152 1 0 1 10      # ViT out-proj: residual ring
  timeout 60 tools/micro/gemm_trace 4796 28672 4096 0 3 0 10      # decoder gate/up + SwiGLU (long K: the reference point)
) > gpurun_out/${T}_gemm_trace.txt 2>&1
cat gpurun_out/${T}_gemm_trace.txt
