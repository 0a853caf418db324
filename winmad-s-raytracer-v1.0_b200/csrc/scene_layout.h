// Host-side builder of the HBM scene layout (see scene_layout.cpp).
#pragma once
#include <string>
#include <vector>
#include "dev_scene.h"

namespace wrt {

struct SceneLayout {
    std::vector<float4> nodes, recs, prims;
    std::vector<DevMaterial> materials;
    std::vector<DevLight> lights;
    int n_nodes = 0;
    int64_t n_recs = 0;
    DevSceneView view;  // pointers left null by build_layout
};

bool build_layout(const wrt_scene_desc* d, SceneLayout& out, std::string& err);
void layout_point_view_at_host(SceneLayout& L);

}  // namespace wrt
