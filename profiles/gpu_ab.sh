for F in 0 4 8; do echo "ARTP_K0_FLAGS=$F"; ARTP_K0_FLAGS=$F bash profiles/gpu_quick.sh; done
