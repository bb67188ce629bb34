#define MRB_TEAM_SCN MRB_PCP
#define MRB_TEAM_TAG pcp
#define MRB_TEAM_N 16
#define MRB_TEAM_WITH_QP
#include "kern_team.inc.h"
