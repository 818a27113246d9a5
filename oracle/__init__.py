"""CPU oracle for the ZkMatrix/ZkVector witness path -- TEST INFRASTRUCTURE ONLY."""
