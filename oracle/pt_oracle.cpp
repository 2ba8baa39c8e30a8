int pt_oracle_placeholder;
