// Stand-in for googletest's gtest_prod.h: the reference uses FRIEND_TEST to let
// its tests read TaskDispatcher's privates (task_dispatcher.h:277-278); we use
// the same hook to give the oracle harness that access.
#pragma once
struct yd_oracle_access;
#define FRIEND_TEST(suite, name) friend struct ::yd_oracle_access
