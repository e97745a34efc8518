// oracle shim (test infrastructure): the vendored nlohmann-json 3.11.3 copy in this image ships only json.hpp.
#pragma once
#include <nlohmann/json.hpp>
