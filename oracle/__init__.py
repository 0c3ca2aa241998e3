"""TEST INFRASTRUCTURE ONLY: CPU restatement + reference-CUDA harness used as the parity checker. Never imported by kintinuous_b200/."""
