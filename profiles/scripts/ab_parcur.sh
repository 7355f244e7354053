# parcur (idim=2) throughput for builds with 3 / 4 / 5 resident CTAs per SM, 2^19 curves
for lib in fitpack_b200/libfitpack_b200_dn3.so fitpack_b200/libfitpack_b200.so fitpack_b200/libfitpack_b200_dn5.so; do echo $lib; FITPACK_B200_LIB=$PWD/$lib python profiles/scripts/curves_only.py parcur 19 2 | tail -1; done
