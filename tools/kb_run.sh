mkdir -p gpurun_out
(
build/kbench k2g 2 64 64 64 16 32 2 10
build/kbench k2s 2 64 64 64 16 32 2 10
build/kbench k2w 2 64 64 64 16 32 2 10
build/kbench k2g 2 32 32 32 32 64 2 10
build/kbench k2s 2 32 32 32 32 64 2 10
build/kbench k2w 2 32 32 32 32 64 2 10
build/kbench k2g 2 8 8 8 128 256 2 10
) > gpurun_out/kb1.log 2>&1
cat gpurun_out/kb1.log
