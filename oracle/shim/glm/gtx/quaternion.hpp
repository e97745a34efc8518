// oracle shim forwarder (test infrastructure) — see glm_subset.hpp
#pragma once
#include <glm/glm_subset.hpp>
