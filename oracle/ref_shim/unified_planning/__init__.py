"""Import stub (test infrastructure) for the three unified_planning names the reference hot path touches
(base_environment.py:3-4, agent_rl.py:1,5). See SURVEY.md Appendix A."""
