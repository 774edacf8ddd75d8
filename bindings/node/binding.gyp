{
  "targets": [{
    "target_name": "gsplat_b200",
    "sources": ["gsplat_napi.cc"],
    "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "../../include"],
    "libraries": ["-L<(module_root_dir)/../../aframe-gaussian-splatting_b200", "-lgsplat_b200"],
    "defines": ["NAPI_DISABLE_CPP_EXCEPTIONS"]
  }]
}
