/* TEST INFRASTRUCTURE.  include/rtiow_b200.h must be usable from plain C (it is what a cgo / Rust-bindgen / JNI
 * binding reads): compiled with `gcc -std=c11 -Wall -Wextra -Werror -pedantic`, linked against the shipped library,
 * run without a GPU (only entry points that do not touch one). */
#include <stdio.h>
#include <string.h>

#include "../../include/rtiow_b200.h"

_Static_assert(sizeof(rtiow_item_t) == 32, "rtiow_item_t");
_Static_assert(sizeof(rtiow_xform_op_t) == 16, "rtiow_xform_op_t");
_Static_assert(sizeof(rtiow_frame_t) == 8, "rtiow_frame_t");
_Static_assert(sizeof(rtiow_material_t) == 32, "rtiow_material_t");
_Static_assert(sizeof(rtiow_texture_t) == 32, "rtiow_texture_t");
_Static_assert(sizeof(rtiow_camera_t) == 84, "rtiow_camera_t");

int main(void) {
    if (rtiow_b200_abi_version() != (int)RTIOW_B200_ABI_VERSION) return 1;
    if (strncmp(rtiow_b200_build_flavour(), "parity", 6) != 0 && strncmp(rtiow_b200_build_flavour(), "fast", 4) != 0) return 2;
    if (rtiow_b200_scene_validate(NULL) != RTIOW_ERR_INVALID_ARG) return 3;
    /* a one-sphere scene, written by hand the way a host language would */
    rtiow_item_t items[2];
    memset(items, 0, sizeof items);
    items[0].a[0] = 0.5f;
    items[0].a_w = RTIOW_ITEM_SPHERE; /* frame 0 */
    items[0].b_w = 0;                 /* material 0 */
    items[1].a_w = RTIOW_ITEM_END;
    rtiow_frame_t world_frame = {0, 0};
    rtiow_texture_t tex;
    memset(&tex, 0, sizeof tex);
    tex.kind = RTIOW_TEX_CONSTANT;
    tex.color[0] = tex.color[1] = tex.color[2] = 0.5f;
    rtiow_material_t mat;
    memset(&mat, 0, sizeof mat);
    mat.kind = RTIOW_MAT_LAMBERTIAN;
    rtiow_scene_desc_t d;
    memset(&d, 0, sizeof d);
    d.abi_version = RTIOW_B200_ABI_VERSION;
    d.n_items = 2; d.items = items;
    d.n_frames = 1; d.frames = &world_frame;
    d.n_materials = 1; d.materials = &mat;
    d.n_textures = 1; d.textures = &tex;
    d.background_kind = RTIOW_BG_SKY_GRADIENT;
    if (rtiow_b200_scene_validate(&d) != RTIOW_OK) { fprintf(stderr, "%s\n", rtiow_b200_last_error()); return 4; }
    items[0].b_w = 7; /* material out of range */
    if (rtiow_b200_scene_validate(&d) != RTIOW_ERR_INVALID_SCENE || strstr(rtiow_b200_last_error(), "material") == NULL) return 5;
    puts("abi_smoke ok");
    return 0;
}
