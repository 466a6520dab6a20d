// prl_host.h -- host-side objects behind the opaque handles of include/prl.h.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "prl_world.h"

struct prl_world {
    float player_position[2];
    std::vector<prl_object> objects;       // World.objects (common.rs:22)
    std::vector<int32_t> object_body;      // from_world's Vec<Option<RigidBodyHandle>> as indices, -1 = None
    prl::DevWorld dev;                     // physics-space tables of Environment::from_world
};

namespace prl {

void set_error(const char* fmt, ...);

// Environment::new + add_object over world.objects (common.rs:70-172) -> static tables.
// Returns PRL_OK or PRL_ERR_CAPACITY.
int build_dev_world(prl_world* w);
// serde_json schema of `World` (common.rs:19-49)
int parse_world_json(const char* s, size_t len, prl_world* out);
std::string world_to_json(const prl_world* w);

}  // namespace prl
