// TEST INFRASTRUCTURE ONLY (oracle/). Shadows the reference's include/freeimpala/learner.h: Learner is fi_host::Learner
// (same constructor order Learner(p,B,S,M,r,c,l,m,T), start, stop, getSharedBuffers, getModelManager; learner.h:100-207).
#pragma once
#include "freeimpala/data_structures.h"
using Learner = fi_host::Learner;
