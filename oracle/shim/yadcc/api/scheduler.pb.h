// Plain-struct stand-in for protoc output of yadcc/api/scheduler.proto: the
// two enums (:38-62) and RunningTask (:233-238), which is all that
// task_dispatcher.{h,cc} and running_task_bookkeeper.{h,cc} use.
#pragma once
#include <cstdint>
#include <string>
#include "yadcc/api/env_desc.pb.h"
namespace yadcc::scheduler {
enum ServantPriority : int {
  SERVANT_PRIORITY_UNKNOWN = 0,
  SERVANT_PRIORITY_DEDICATED = 1,
  SERVANT_PRIORITY_USER = 2,
};
inline bool ServantPriority_IsValid(int v) { return v >= 0 && v <= 2; }
inline const std::string& ServantPriority_Name(ServantPriority v) {
  static const std::string n[] = {"SERVANT_PRIORITY_UNKNOWN", "SERVANT_PRIORITY_DEDICATED",
                                  "SERVANT_PRIORITY_USER", ""};
  return n[(v >= 0 && v <= 2) ? v : 3];
}
enum NotAcceptingTaskReason : int {
  NOT_ACCEPTING_TASK_REASON_UNKNOWN = 0,
  NOT_ACCEPTING_TASK_REASON_USER_INSTRUCTED = 1,
  NOT_ACCEPTING_TASK_REASON_POOR_MACHINE = 2,
  NOT_ACCEPTING_TASK_REASON_CGROUPS_PRESENT = 3,
  NOT_ACCEPTING_TASK_REASON_BEHIND_NAT = 4,
  NOT_ACCEPTING_TASK_REASON_NOT_VERIFIED = 100,
};
inline const std::string& NotAcceptingTaskReason_Name(NotAcceptingTaskReason v) {
  static const std::string n[] = {
      "NOT_ACCEPTING_TASK_REASON_UNKNOWN",         "NOT_ACCEPTING_TASK_REASON_USER_INSTRUCTED",
      "NOT_ACCEPTING_TASK_REASON_POOR_MACHINE",    "NOT_ACCEPTING_TASK_REASON_CGROUPS_PRESENT",
      "NOT_ACCEPTING_TASK_REASON_BEHIND_NAT",      "NOT_ACCEPTING_TASK_REASON_NOT_VERIFIED", ""};
  return n[(v >= 0 && v <= 4) ? v : (v == 100 ? 5 : 6)];
}
class RunningTask {
 public:
  std::uint64_t servant_task_id() const { return servant_task_id_; }
  std::uint64_t task_grant_id() const { return task_grant_id_; }
  const std::string& servant_location() const { return servant_location_; }
  const std::string& task_digest() const { return task_digest_; }
  void set_servant_task_id(std::uint64_t v) { servant_task_id_ = v; }
  void set_task_grant_id(std::uint64_t v) { task_grant_id_ = v; }
  void set_servant_location(std::string v) { servant_location_ = std::move(v); }
  void set_task_digest(std::string v) { task_digest_ = std::move(v); }

 private:
  std::uint64_t servant_task_id_ = 0, task_grant_id_ = 0;
  std::string servant_location_, task_digest_;
};
}  // namespace yadcc::scheduler
