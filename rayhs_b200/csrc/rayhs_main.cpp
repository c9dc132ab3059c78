// rayhs_main.cpp — the `rayhs [-oFILE | --output[=FILE]] scene.json` driver (RayHs.hs:204-234) over the C ABI.
// C++ stand-in for the Haskell `main`, which stays the real front end where GHC exists
// (INTEGRATION.md).  Same option semantics as GetOpt RequireOrder with an optional-argument
// `-o` (the file name must be glued: -ocornell.ppm), same default (out.ppm), same progress lines.
// Extra environment knobs (not in the reference): RAYHS_SPP (samples per pixel, default 1 =
// rayTrace; >1 = distributedRayTrace with the RayHs.hs:239 seed 24), RAYHS_WIDTH / RAYHS_HEIGHT,
// RAYHS_GPUS (render on that many GPUs of this box through rh_multi_render).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rayhs_b200.h"

static int fail(const char* what) {
  fprintf(stderr, "rayhs: %s: %s\n", what, rh_last_error());
  return 1;
}

int main(int argc, char** argv) {
  std::string out_file = "out.ppm";  // RayHs.hs:213, 220-222
  int i = 1;
  for (; i < argc; i++) {  // RequireOrder: options must precede the scene file
    const char* a = argv[i];
    if (a[0] != '-' || !a[1]) break;
    if (!strncmp(a, "--output", 8)) out_file = (a[8] == '=') ? a + 9 : "out.ppm";
    else if (a[1] == 'o') out_file = a[2] ? a + 2 : "out.ppm";
    else {
      fprintf(stderr, "unrecognized option `%s'\nUsage: main [OPTION...]\n  -o[FILE]  --output[=FILE]  Specify the ouput filename.\n", a);
      return 1;
    }
  }
  if (i >= argc) {
    fprintf(stderr, "rayhs: No input.\n");  // RayHs.hs:233
    return 1;
  }
  const char* fn = argv[i];
  printf("Loading scene from %s...\n", fn);  // RayHs.hs:223
  rh_loaded* loaded = nullptr;
  if (rh_load_json(fn, "", &loaded)) return fail("Failed to read scene");  // RayHs.hs:226
  printf("Rendering...\n");  // RayHs.hs:228
  rh_flat_scene* flat = nullptr;
  if (rh_flatten(rh_loaded_raw(loaded), &flat)) return fail("flatten");
  const int n_gpus = getenv("RAYHS_GPUS") ? atoi(getenv("RAYHS_GPUS")) : 1;
  rh_scene* scene = nullptr;
  rh_multi_scene* mscene = nullptr;
  if (n_gpus > 1) {
    if (rh_multi_init(n_gpus)) return fail("rh_multi_init");
    if (rh_multi_scene_create(rh_flat_desc(flat), &mscene)) return fail("rh_multi_scene_create");
  } else {
    if (rh_init(-1)) return fail("rh_init");
    if (rh_scene_create(rh_flat_desc(flat), &scene)) return fail("rh_scene_create");
  }
  rh_render_opts o;
  memset(&o, 0, sizeof o);
  rh_loaded_size(loaded, &o.width, &o.height, &o.max_depth);
  if (const char* e = getenv("RAYHS_WIDTH")) o.width = atoi(e);
  if (const char* e = getenv("RAYHS_HEIGHT")) o.height = atoi(e);
  o.spp = 1;
  if (const char* e = getenv("RAYHS_SPP")) o.spp = atoi(e) > 0 ? atoi(e) : 1;
  o.shard_count = 1;
  std::vector<double> offsets;
  if (o.spp > 1) {  // distributedRayTrace (RayHs.hs:190-195, 239)
    offsets.resize((size_t)o.width * o.height * o.spp * 2);
    rh_sample_offsets_f64(24, (uint64_t)o.width * o.height, o.spp, offsets.data());
    o.offset_mode = RH_OFFSETS_F64;
    o.offsets = offsets.data();
  }
  std::vector<uint8_t> rgb((size_t)o.width * o.height * 3);
  rh_stats st;
  if (n_gpus > 1) {
    if (rh_multi_render(mscene, rh_loaded_camera(loaded), &o, rgb.data(), &st)) return fail("rh_multi_render");
  } else if (rh_render(scene, rh_loaded_camera(loaded), &o, rgb.data(), nullptr, &st)) {
    return fail("rh_render");
  }
  if (rh_write_ppm(out_file.c_str(), rgb.data(), o.width, o.height)) return fail("writePPM");
  printf("Done! Output written to %s\n", out_file.c_str());  // RayHs.hs:232
  if (getenv("RAYHS_STATS")) {
    const unsigned long long rays = st.rays_primary + st.rays_reflect + st.rays_probe + st.rays_exit + st.rays_shadow;
    fprintf(stderr, "%.3f ms on the GPU, %llu rays (%.1f Mrays/s), %u launches\n", st.ms_total, rays, rays / st.ms_total / 1e3,
            st.kernel_launches);
  }
  if (n_gpus > 1) {
    rh_multi_scene_destroy(mscene);
    rh_multi_shutdown();
  } else {
    rh_scene_destroy(scene);
    rh_shutdown();
  }
  rh_flat_destroy(flat);
  rh_loaded_destroy(loaded);
  return 0;
}
