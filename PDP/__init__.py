"""placeholder -- filled in below"""
