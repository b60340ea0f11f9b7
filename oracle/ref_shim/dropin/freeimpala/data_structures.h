// TEST INFRASTRUCTURE ONLY (oracle/). The alias header INTEGRATION.md describes, in the form the drop-in proof needs:
// it shadows the reference's include/freeimpala/data_structures.h on the include path, so that the UNMODIFIED
// include/freeimpala/agent.h (which includes "freeimpala/data_structures.h") compiles against freeimpala_b200/host/fi_host.hpp.
//
// The actor-private pieces -- ELEMENT_SIZE, MessageTag, BufferEntry, Buffer (data_structures.h:21-35, 160-188) -- are the
// reference's own definitions: its header is included, where it lies, inside a namespace (every standard header it pulls in
// is included first, so their include guards keep them out of that namespace) and those four names are re-exported.
// SharedBuffer / Model / ModelManager are the fi_host classes. Nothing of the reference is copied into the repo.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <queue>
#include <string>
#include <thread>
#include <vector>
#include <spdlog/spdlog.h>

namespace fi_reference {
#include FI_REF_DATA_STRUCTURES_H   // "/root/reference/include/freeimpala/data_structures.h"
}
using fi_reference::Buffer;
using fi_reference::BufferEntry;
using fi_reference::ELEMENT_SIZE;
using fi_reference::MessageTag;

// the reference's MetricsTracker, unmodified: fi_host::Learner::trainModel makes the two calls of learner.h:34,48 on it
#include "freeimpala/metrics_tracker.h"
#define FI_HOST_METRICS_TRACKER 1
#include <fi_host.hpp>
using SharedBuffer = fi_host::SharedBuffer;   // write / try_write / readBatch / setDraining / getFilledCount
using Model = fi_host::Model;                 // getVersion / getData / createCopy / update
using ModelManager = fi_host::ModelManager;   // getModel / getLatestVersion / waitForModelUpdate / saveModel / loadModels
