// TEST INFRASTRUCTURE.  etcdv3::Transaction lives in etcd/SyncClient.hpp of this stub set.
#pragma once
#include "etcd/SyncClient.hpp"
