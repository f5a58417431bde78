for V in lb8 lb10; do cp variants_tmp/libartp_$V.so art_planner_b200/libartp.so; echo "variant $V"; ARTP_SKIP_BUILD=1 bash profiles/gpu_quick.sh; done
