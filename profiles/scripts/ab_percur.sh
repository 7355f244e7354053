# percur throughput for several builds (resident CTAs per SM), 2^20 curves
for lib in fitpack_b200/libfitpack_b200_pb4.so fitpack_b200/libfitpack_b200.so fitpack_b200/libfitpack_b200_pb6.so; do echo $lib; FITPACK_B200_LIB=$PWD/$lib python profiles/scripts/curves_only.py percur 20 2 | tail -1; done
