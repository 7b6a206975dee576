#!/bin/bash
# session 3, call 9: chunk count of the pipelined upload (packed host form)
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/trace_upload.py packed 1300000 8,12,16,24,32 > gpurun_out/upload_chunks.log 2>&1; grep -E "pack_csc|total|---- upload" gpurun_out/upload_chunks.log | paste - - | cut -c1-120
