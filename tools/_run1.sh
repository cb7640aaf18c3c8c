cd /root/repo
DCLIP_B200_LIB=build/prof/libdclip_b200_prof.so timeout 120 python tools/prof_small_loss.py 2>&1 | tail -16
