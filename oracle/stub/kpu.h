/* Empty stand-in for the K210 SDK's kpu.h.
 * /root/reference/yolo3_frame_test_public/region_layer.h:4 includes it but
 * region_layer.c uses no symbol from it, so the unmodified reference source
 * compiles on a host with this stub on the include path (oracle/Makefile). */
