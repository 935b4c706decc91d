// test scaffolding: see tests/stubs/ocs2_stub.hpp
#pragma once
#include "../../ocs2_stub.hpp"
