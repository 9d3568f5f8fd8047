/* c_abi_consumer.c -- include/splat_b200.h is a C header: this file is compiled with gcc -std=c99 -pedantic -Wall -Werror
 * and linked against libsplat_b200.so (what a cgo / JNI / N-API binding sees).  No GPU needed: it exercises the host side
 * of the ABI and the error channel. */
#include <stdio.h>
#include <string.h>

#include "splat_b200.h"

int main(int argc, char** argv)
{
    float origin[3] = {0.0f, 0.0f, 0.0f};
    size_t n, stride;
    int type = -1;
    size_t size = 0, offset = 0;
    spl_cloud* cloud;
    if (spl_version() < 100) return 1;
    cloud = spl_cloud_create(1, 1);
    if (!cloud) return 2;
    if (spl_cloud_init_debug_cloud(cloud) != SPL_OK) return 3;
    n = spl_cloud_num_gaussians(cloud);
    stride = spl_cloud_stride(cloud);
    if (n != 16 || stride != 244 || spl_cloud_total_size(cloud) != n * stride || !spl_cloud_raw_data(cloud)) return 4;
    if (spl_cloud_attrib(cloud, SPL_ATTR_POS_WITH_ALPHA, &type, &size, &offset) != SPL_OK || type != 7 || size != 4 || offset != 0) return 5;
    if (spl_cloud_prune_splats(cloud, origin, 4) != SPL_OK || spl_cloud_num_gaussians(cloud) != 4) return 6;
    if (argc > 1 && spl_cloud_export_ply(cloud, argv[1]) != SPL_OK) return 7;
    if (spl_cloud_import_ply(cloud, "/nonexistent/file.ply") == SPL_OK || strlen(spl_last_error()) == 0) return 8;
    if (spl_cloud_import_ply(NULL, "x") != SPL_ERR_ARG) return 9;
    spl_cloud_destroy(cloud);
    printf("c abi ok: %d device(s)\n", spl_device_count());
    return 0;
}
