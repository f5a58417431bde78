for F in 0 2 4 6; do echo "ARTP_K1_FLAGS=$F"; ARTP_K1_FLAGS=$F bash profiles/gpu_quick.sh; done
